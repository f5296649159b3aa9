"""Integer restatement of the motion-mask pyramid.  TEST INFRASTRUCTURE.

Follows ``ImageProcessor.preprocess_mov_mask`` (src/dataset/image_processor.py:311-333) whose
transforms (``:75-102``) are torchvision ``Resize`` -> Pillow ``Image.resize(BILINEAR)`` on 8-bit
``L`` images, followed by ``ToTensor`` (uint8/255 in fp32).  Pillow (pinned 9.5.0,
requirements.txt:138) resamples in two passes -- horizontal then vertical -- with 22-bit fixed
point coefficients and a uint8 intermediate.  Pinned bit-exactly against the live
torchvision+Pillow in this image by tests/test_mask_pyramid.py.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Pillow: 22


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs() for the triangle filter (support 1.0)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            w[x] = max(0.0, 1.0 - abs(v))
        tot = w[:xmax].sum()
        if tot != 0.0:
            w[:xmax] /= tot
        for x in range(ksize):
            c = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(c - 0.5) if c < 0 else int(c + 0.5)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img (..., H, W) uint8 -> (..., out_h, out_w) uint8, Pillow semantics."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape[-2:]
    cur = img
    if out_w != w:
        bounds, kk = bilinear_coeffs(w, out_w)
        out = np.empty(cur.shape[:-1] + (out_w,), dtype=np.uint8)
        for xx in range(out_w):
            xmin, n = bounds[xx]
            acc = (cur[..., xmin:xmin + n].astype(np.int64) * kk[xx, :n]).sum(-1) + (1 << (PRECISION_BITS - 1))
            out[..., xx] = _clip8(acc)
        cur = out
    if out_h != h:
        bounds, kk = bilinear_coeffs(h, out_h)
        out = np.empty(cur.shape[:-2] + (out_h, cur.shape[-1]), dtype=np.uint8)
        for yy in range(out_h):
            ymin, n = bounds[yy]
            acc = (cur[..., ymin:ymin + n, :].astype(np.int64) * kk[yy, :n, None]).sum(-2) + (1 << (PRECISION_BITS - 1))
            out[..., yy, :] = _clip8(acc)
        cur = out
    return cur


def mask_pyramid_u8(masks: np.ndarray, image_size: int = 512):
    """masks (L, Hs, Ws) uint8 -> 4 levels of uint8 (L, S_k, S_k), S_k = image_size // (8 << k)."""
    out = []
    for k in range(4):
        s = image_size // (8 << k)
        out.append(resize_bilinear_u8(masks, s, s))
    return out


def preprocess_mov_mask(face_u8: np.ndarray, lips_u8: np.ndarray, image_size: int = 512):
    """-> (face levels, lips levels): lists of 4 float32 arrays (L, S_k*S_k) = uint8 / 255."""
    face = [(m.astype(np.float32) / np.float32(255.0)).reshape(m.shape[0], -1)
            for m in mask_pyramid_u8(face_u8, image_size)]
    lips = [(m.astype(np.float32) / np.float32(255.0)).reshape(m.shape[0], -1)
            for m in mask_pyramid_u8(lips_u8, image_size)]
    return face, lips


def full_mask_from_lips(lips_levels):
    """scripts/audio2vid.py:470-476: the loop that survives is ``full = 1.0 + lips``."""
    return [np.float32(1.0) + m for m in lips_levels]
