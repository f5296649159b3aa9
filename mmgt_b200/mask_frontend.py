"""Motion-mask front-end of the inference scripts: raw mask video frames -> the three lists of four
``(L, T_l)`` tensors ``Pose2VideoPipeline.__call__`` takes as ``pixel_values_{full,face,lip}_mask``.

  frames --blur_mask--> 64x64 8-bit "L" images --MaskPyramid (bit-exact Pillow bilinear, CUDA)--> 4 levels --recipe--> full

* ``blur_mask`` is the scripts' own helper (scripts/pose2vid.py:94-114 == src/utils/util.py:19-39): ``cv2.resize`` to 64x64,
  ``cv2.GaussianBlur``, ``cv2.normalize(0, 255, NORM_MINMAX)``.  It runs once per frame on the host with the same OpenCV
  calls the reference makes (OpenCV is the reference's dependency for it), so its output is identical by construction.
* The pyramid is ``mmgt_b200.image_processor.MaskPyramid`` (``ImageProcessor.preprocess_mov_mask``,
  src/dataset/image_processor.py:311-333) on the GPU.
* The "full" mask recipe.  scripts/audio2vid.py:470-476 builds ``full = 1 + lips`` per level (the ``1 - face`` it computes
  first is overwritten).  scripts/pose2vid.py:262-271 intends ``clamp(1 - face + lips + hands, 0, 1)`` but indexes the
  list of 4 LEVELS with the FRAME index and adds a (L, 4096) tensor to a (1, 64, 64) one, so it raises at i = 0 (SURVEY fact
  10).  ``recipe="pose2vid"`` implements what that loop means -- per level l: ``clamp(1 - face_l + lips_l + hands_l, 0, 1)``
  with the hands mask taken through the same pyramid (the script only ever resized it to 64x64 = level 0 at 512x512).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

FACE_KERNEL, LIPS_KERNEL, HANDS_KERNEL = (31, 31), (21, 21), (21, 21)      # scripts/pose2vid.py:247-251


def blur_mask(mask, resize_dim=(64, 64), kernel_size=(51, 51)):
    """src/utils/util.py:19-39, unchanged semantics (None in, None out)."""
    if mask is None:
        return None
    import cv2
    resized = cv2.resize(mask, resize_dim)
    blurred = cv2.GaussianBlur(resized, kernel_size, 0)
    return cv2.normalize(blurred, None, 0, 255, cv2.NORM_MINMAX)


def prepare_mask_frames(frames: Sequence, kernel_size, length: Optional[int] = None) -> List["object"]:
    """scripts/pose2vid.py:240-246 ``_prep_mask_list``: PIL / array frames -> blurred 64x64 PIL "L" images."""
    from PIL import Image
    out = []
    for img in list(frames)[:length]:
        proc = blur_mask(np.array(img), resize_dim=(64, 64), kernel_size=kernel_size)
        out.append(Image.fromarray(proc.astype(np.uint8)).convert("L"))
    return out


def full_mask_levels(face: Sequence[torch.Tensor], lips: Sequence[torch.Tensor], hands: Optional[Sequence[torch.Tensor]],
                     recipe: str) -> List[torch.Tensor]:
    if recipe == "audio2vid":                      # scripts/audio2vid.py:470-476
        return [1.0 + lp.float() for lp in lips]
    if recipe == "pose2vid":                       # the intent of scripts/pose2vid.py:262-271, per pyramid level
        out = []
        for lvl, (fc, lp) in enumerate(zip(face, lips)):
            full = (1.0 - fc.float()) + lp.float()
            if hands is not None:
                full = full + hands[lvl].float()
            out.append(torch.clamp(full, 0.0, 1.0))
        return out
    raise ValueError(f"unknown full-mask recipe {recipe!r} (audio2vid | pose2vid)")


def motion_masks(face_frames: Sequence, lips_frames: Sequence, hands_frames: Optional[Sequence] = None, image_size: int = 512,
                 recipe: str = "audio2vid", length: Optional[int] = None, pyramid=None,
                 device="cuda") -> Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    """Raw mask frames -> (pixel_values_full_mask, pixel_values_face_mask, pixel_values_lip_mask), each a list of four
    ``(L, (image_size / (8 << k))**2)`` float32 tensors, as the scripts hand them to the pipeline.  ``image_size`` is the
    script's ``config.data.source_image.width`` (it must equal the generated width).
    ``pyramid``: object with ``levels(list_of_L_images) -> 4 tensors`` (default: the CUDA MaskPyramid)."""
    if pyramid is None:
        from .image_processor import MaskPyramid
        pyramid = MaskPyramid(image_size, device)
    face_l = prepare_mask_frames(face_frames, FACE_KERNEL, length)
    lips_l = prepare_mask_frames(lips_frames, LIPS_KERNEL, length)
    if len(face_l) != len(lips_l) or not face_l:
        raise ValueError(f"need the same, non-zero number of face and lips mask frames (got {len(face_l)} / {len(lips_l)})")
    face = pyramid.levels(face_l)
    lips = pyramid.levels(lips_l)
    hands = None
    if hands_frames is not None:
        hands_l = prepare_mask_frames(hands_frames, HANDS_KERNEL, length)
        if len(hands_l) != len(face_l):
            raise ValueError("hands mask frames must match the face mask frames")
        hands = pyramid.levels(hands_l)
    return full_mask_levels(face, lips, hands, recipe), face, lips
