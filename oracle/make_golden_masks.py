"""Real motion-mask fixture ("also run the bundled real fixture for masks").  TEST INFRASTRUCTURE.

Runs ONLY where /root/reference exists.  Reads the first frames of config/cases/oliver#103842_slice18_{face,lips}_mask.mp4
and applies the scripts' own front-end -- blur_mask (scripts/pose2vid.py:94-114 = src/utils/util.py:19-39: cv2.resize to
64 x 64, cv2.GaussianBlur, cv2.normalize 0..255 MINMAX) with the (31, 31) / (21, 21) kernels of scripts/pose2vid.py:247-248,
then PIL convert("L") (:244) -- and stores the resulting 64 x 64 uint8 images, i.e. exactly what
ImageProcessor.preprocess_mov_mask receives.  The pyramid itself is then checked bit-exactly: oracle vs live
torchvision + Pillow on CPU, CUDA kernel vs oracle on the GPU.

Usage:  python -m oracle.make_golden_masks
"""
import os

import numpy as np

from . import reference_loader as RL

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FRAMES = (0, 13, 26, 39, 52, 79)


def front_end(rgb: np.ndarray, ksize) -> np.ndarray:
    import cv2
    from PIL import Image
    resized = cv2.resize(rgb, (64, 64))
    blurred = cv2.GaussianBlur(resized, ksize, 0)
    normalized = cv2.normalize(blurred, None, 0, 255, cv2.NORM_MINMAX)
    return np.array(Image.fromarray(normalized.astype(np.uint8)).convert("L"))


def read_frames(path, which):
    import cv2
    cap = cv2.VideoCapture(path)
    out, i = {}, 0
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        if i in which:
            out[i] = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        i += 1
    return [out[i] for i in which]


def main():
    base = os.path.join(RL.REFERENCE_ROOT, "config", "cases", "oliver#103842_slice18_{}_mask.mp4")
    face = np.stack([front_end(f, (31, 31)) for f in read_frames(base.format("face"), FRAMES)])
    lips = np.stack([front_end(f, (21, 21)) for f in read_frames(base.format("lips"), FRAMES)])
    assert face.shape == lips.shape == (len(FRAMES), 64, 64) and face.dtype == np.uint8
    np.savez_compressed(os.path.join(GOLD, "real_masks.npz"), face=face, lips=lips)
    print("wrote real_masks.npz", face.shape, "face levels", np.unique(face).size, "lips levels", np.unique(lips).size)


if __name__ == "__main__":
    main()
