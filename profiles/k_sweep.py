#!/usr/bin/env python
"""K sweep of the small-K GEMMs (M = 24 x 4096 rows): does the tile period grow with the main loop (epilogue and MMA
serialised) or stay flat (overlapped, epilogue-bound)?  python profiles/k_sweep.py  -> one line per (kind, K)."""
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
    M = 24 * 4096
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for kind, N, geglu, res in (("plain", 960, False, False), ("geglu", 2560, True, False), ("res", 320, False, True),
                                ("plain320", 320, False, False), ("plain640", 640, False, False)):
        for K in (64, 128, 192, 256, 320, 448, 640, 1280):
            A = (torch.randn(M, K, generator=g)).to(device=dev, dtype=torch.bfloat16)
            W = (torch.randn(N, K, generator=g) * K ** -0.5).to(device=dev, dtype=torch.bfloat16)
            bias = torch.randn(N, generator=g).to(dev)
            n_out = N // 2 if geglu else N
            r = torch.randn(M, n_out, generator=g).to(device=dev, dtype=torch.bfloat16) if res else None
            out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
            gb = eng.geglu_block(N) if geglu else 0
            fn = lambda: eng.gemm(A, W, bias=bias, residual=r, geglu_block=gb, out=out)  # noqa: E731
            fn(); fn()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            us = ts[2] * 1e3
            print(f"{kind:9s} N={N:5d} K={K:5d}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
