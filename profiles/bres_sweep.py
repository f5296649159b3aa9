#!/usr/bin/env python
"""Weight-stationary (resident B) vs streaming GEMM kernel on the K = 640 / 1280-row-tile shapes of levels 1-2:
python profiles/bres_sweep.py -> one line per (M, N, K, residual, mode)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mmgt_b200.kernels import get_engine  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    eng = get_engine(dev, torch.bfloat16)
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    shapes = [(24576, 640, 640, True), (24576, 640, 640, False), (24576, 1920, 640, False), (98304, 640, 640, True),
              (24576, 640, 2560, True), (98304, 320, 320, True), (98304, 960, 320, False), (24576, 5120, 640, "geglu"),
              (6144, 1280, 1280, True), (6144, 3840, 1280, False), (12288, 640, 640, True), (12288, 1920, 640, False)]
    for M, N, K, kind in shapes:
        A = torch.randn(M, K, generator=g).to(device=dev, dtype=torch.bfloat16)
        W = (torch.randn(N, K, generator=g) * K ** -0.5).to(device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, generator=g).to(dev)
        geglu = kind == "geglu"
        n_out = N // 2 if geglu else N
        r = torch.randn(M, n_out, generator=g).to(device=dev, dtype=torch.bfloat16) if kind is True else None
        out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        gb = eng.geglu_block(N) if geglu else 0
        fn = lambda: eng.gemm(A, W, bias=bias, residual=r, geglu_block=gb, out=out)  # noqa: E731
        for bres in (True, False, "lane-residual"):
            eng.ctx.set_resident_weights(bres is True)
            eng.ctx.set_residual_mma(bres != "lane-residual")
            fn(); fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            us = ts[3] * 1e3
            print(f"M={M:6d} N={N:5d} K={K:5d} {str(kind):6s} resident={str(bres):14s}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s",
                  flush=True)
        eng.ctx.set_resident_weights(True)
        eng.ctx.set_residual_mma(True)


if __name__ == "__main__":
    main()
