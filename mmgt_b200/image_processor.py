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


class VaeImageProcessor:
    """PIL / tensor -> (B, 3, H, W) float32, the subset of diffusers==0.24.0 ``VaeImageProcessor.preprocess`` that
    ``Pose2VideoPipeline`` uses (pipeline_pose2vid_long.py:72-79,424-437): optional RGB conversion, Lanczos resize to
    (height, width) rounded down to multiples of ``vae_scale_factor``, ``uint8 / 255``, optional ``2x - 1``.  Host-side
    PIL / torch code (one-shot conditioning, not on the timed path)."""

    def __init__(self, vae_scale_factor: int = 8, do_resize: bool = True, do_normalize: bool = True,
                 do_convert_rgb: bool = False):
        self.vae_scale_factor, self.do_resize = vae_scale_factor, do_resize
        self.do_normalize, self.do_convert_rgb = do_normalize, do_convert_rgb

    def preprocess(self, image, height: int = None, width: int = None) -> torch.Tensor:
        from PIL import Image
        if isinstance(image, (Image.Image, np.ndarray, torch.Tensor)):
            image = [image]
        if isinstance(image[0], Image.Image):
            if self.do_convert_rgb:
                image = [i.convert("RGB") for i in image]
            if self.do_resize:
                height = image[0].height if height is None else height
                width = image[0].width if width is None else width
                width, height = (x - x % self.vae_scale_factor for x in (width, height))
                image = [i.resize((width, height), resample=Image.LANCZOS) for i in image]
            arr = np.stack([np.array(i).astype(np.float32) / 255.0 for i in image], axis=0)
            if arr.ndim == 3:
                arr = arr[..., None]
            out = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        elif isinstance(image[0], np.ndarray):
            arr = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            if arr.ndim == 3:
                arr = arr[..., None]
            out = torch.from_numpy(arr.transpose(0, 3, 1, 2)).float()
        elif torch.is_tensor(image[0]):
            out = torch.cat(image, dim=0) if image[0].dim() == 4 else torch.stack(image, dim=0)
            out = out.float()
        else:
            raise ValueError(f"unsupported image type {type(image[0])}")
        if self.do_normalize and float(out.min()) >= 0:
            out = 2.0 * out - 1.0
        return out
