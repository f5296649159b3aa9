"""Motion-mask pyramid on the GPU (host mirror of ImageProcessor.preprocess_mov_mask,
src/dataset/image_processor.py:75-102,311-333).

Input masks are the 64x64 8-bit ``L`` images the scripts build with ``blur_mask`` (scripts/pose2vid.py:217-263);
output is the reference's list of four ``(L, S_k*S_k)`` float32 tensors with S_k = image_size / (8 << k),
bit-identical to torchvision ``Resize`` (Pillow bilinear) + ``ToTensor``.
"""
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

from .kernels import get_engine


def _to_u8_stack(masks, device) -> torch.Tensor:
    if torch.is_tensor(masks):
        t = masks
    else:
        t = torch.from_numpy(np.stack([np.asarray(m, dtype=np.uint8) for m in masks]))
    if t.dtype != torch.uint8 or t.dim() != 3:
        raise ValueError("masks must be (L, H, W) uint8 (8-bit 'L' images)")
    return t.to(device).contiguous()


class MaskPyramid:
    def __init__(self, image_size: int = 512, device: Union[str, torch.device] = "cuda"):
        self.image_size = image_size
        self.device = torch.device(device)
        self.sizes = [image_size // (8 << k) for k in range(4)]

    def levels(self, masks, offset: float = 0.0) -> List[torch.Tensor]:
        eng = get_engine(self.device, torch.float32)
        src = _to_u8_stack(masks, eng.device)
        return [eng.mask_resize(src, s, offset=offset) for s in self.sizes]

    def preprocess_mov_mask(self, face_masks_list: Sequence, lips_masks_list: Sequence, face_region_ratio: float = 0.0,
                            clip_length: int = None) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """-> (pixel_values_face_mask, pixel_values_lips_mask) as in image_processor.py:311-333."""
        assert face_masks_list[0] is not None, "Fail to load face mask."
        assert lips_masks_list[0] is not None, "Fail to load lip mask."
        return self.levels(face_masks_list), self.levels(lips_masks_list)

    def full_mask_from_lips(self, lips_masks_list: Sequence) -> List[torch.Tensor]:
        """``full = 1.0 + lips`` per level (the recipe that survives in scripts/audio2vid.py:470-476)."""
        return self.levels(lips_masks_list, offset=1.0)
