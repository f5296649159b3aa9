"""GPU: frame-sharded execution (SURVEY.md section 8e level 3) -- the row exchange fused into the GEMM epilogue,
the stand-alone exchange copy, the flag barrier, and a whole UNet window split over k shards.

Single-GPU cases emulate the k shards inside one process (FrameShardGroup.emulate: the "peer" buffers are
ordinary allocations on the same device, one thread + stream per shard), so the exact kernels of the multi-GPU
path run under the driver's 1-GPU `pytest -m gpu`.  The real thing -- one process per GPU, CUDA IPC, NVLink peer
stores -- is tests/multigpu_frame_shard.py, launched here through torchrun when two GPUs are visible."""
import os
import subprocess
import sys
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import TINY, attach_banks, build_cuda_unet, rel_l2, synthetic_state_dict, to_dev  # noqa: E402
from oracle.synthetic import make_banks, make_inputs, window_inputs  # noqa: E402
from oracle.unet3d import UNetSpec  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _engine(dev, dtype):
    from mmgt_b200.kernels import Engine
    return Engine(dev, dtype)


def _relayout(full, direction, k, B, F, T):
    """torch statement of the exchange: list (per shard) of source rows and of expected received rows."""
    Fl, Tc, C = F // k, T // k, full.shape[-1]
    x = full.view(B, F, T, C)
    frame = [x[:, s * Fl:(s + 1) * Fl].reshape(-1, C) for s in range(k)]
    token = [x[:, :, s * Tc:(s + 1) * Tc].reshape(-1, C) for s in range(k)]
    return (frame, token) if direction == 1 else (token, frame)


@pytest.mark.parametrize("k,B,F,T,K,N", [(2, 2, 12, 256, 64, 64), (2, 1, 12, 64, 320, 320), (4, 2, 12, 64, 128, 256),
                                         (4, 1, 4, 16, 1280, 1280), (2, 2, 6, 1024, 320, 640)])
@pytest.mark.parametrize("direction", [1, 2])
def test_gemm_epilogue_row_exchange(dev, k, B, F, T, K, N, direction):
    """The GEMM that produces the rows delivers them: bit-identical to GEMM-then-re-layout, both directions, with the
    residual / bias epilogue, and the stand-alone copy kernel agrees."""
    from mmgt_b200.frame_shard import FrameShardGroup
    eng = _engine(dev, torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    A_full = torch.randn(B * F * T, K, generator=g).to(dev, torch.bfloat16)
    R_full = torch.randn(B * F * T, N, generator=g).to(dev, torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    src_A, _ = _relayout(A_full, direction, k, B, F, T)
    src_R, _ = _relayout(R_full, direction, k, B, F, T)
    groups = FrameShardGroup.emulate(eng, k, B * (F // k) * T * N * 2)
    try:
        # The exchange epilogue adds bias and residual per lane, (acc + bias) + r; the default small-K GEMM feeds the
        # residual through [R | I] k-blocks on the tensor cores, (acc + r) + bias (ctx flag 13): same value up to the
        # order of two float32 additions.  Bit-identity is asserted against the per-lane form, the default form is held
        # to one bf16 rounding step of it.
        default = [eng.gemm(src_A[s].contiguous(), W, bias=bias, residual=src_R[s].contiguous()) for s in range(k)]
        eng.ctx.set_residual_mma(False)
        plain = [eng.gemm(src_A[s].contiguous(), W, bias=bias, residual=src_R[s].contiguous()) for s in range(k)]
        for d, p_ in zip(default, plain):      # (the flag stays off below: the unfused leg is a plain GEMM + copy kernel)
            err = (d.float() - p_.float()).abs()
            assert bool((err <= 2.0 ** -7 * p_.float().abs() + 2e-6).all()), float(err.max())   # 2e-6: float32 rounding under cancellation
        # what each shard must receive = re-layout of the concatenated plain results
        if direction == 1:
            full = torch.stack([p.view(B, F // k, T, N) for p in plain], dim=1).reshape(B, F, T, N)   # (B, k, Fl, ..)
        else:
            full = torch.stack([p.view(B, F, T // k, N) for p in plain], dim=2)            # (B, F, k, Tc, N)
            full = full.reshape(B, F, T, N)
        _, expect = _relayout(full.reshape(-1, N), direction, k, B, F, T)
        for fused in (True, False):
            eng.unfused_exchange = not fused
            exs = [groups[s].exchange(direction, B, F, T, N) for s in range(k)]
            for ex in exs:
                ex.recv.zero_()
            for s in range(k):
                assert eng.gemm(src_A[s].contiguous(), W, bias=bias, residual=src_R[s].contiguous(), exchange=exs[s]) is None
            torch.cuda.synchronize()
            for s in range(k):
                assert torch.equal(exs[s].recv, expect[s]), f"shard {s} fused={fused}"
    finally:
        eng.ctx.set_residual_mma(True)
        eng.unfused_exchange = False
        groups[0].close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_row_exchange_copy_roundtrip(dev, dtype):
    from mmgt_b200.frame_shard import FrameShardGroup
    eng = _engine(dev, dtype)
    k, B, F, T, C = 4, 2, 8, 36, 64
    full = torch.randn(B * F * T, C, generator=torch.Generator().manual_seed(3)).to(dev, dtype)
    frame, token = _relayout(full, 1, k, B, F, T)
    groups = FrameShardGroup.emulate(eng, k, full.numel() * full.element_size() // k)
    try:
        fwd = [groups[s].exchange(1, B, F, T, C) for s in range(k)]
        for s in range(k):
            eng.row_exchange_copy(frame[s].contiguous(), fwd[s])
        torch.cuda.synchronize()
        for s in range(k):
            assert torch.equal(fwd[s].recv, token[s])
        back = [groups[s].exchange(2, B, F, T, C) for s in range(k)]
        for s in range(k):
            eng.row_exchange_copy(fwd[s].recv, back[s])
        torch.cuda.synchronize()
        for s in range(k):
            assert torch.equal(back[s].recv, frame[s])
    finally:
        groups[0].close()


def test_peer_barrier_protocol_and_timeout(dev):
    """Two shards on two streams pass 5 barriers; a shard whose peer never arrives reports a timeout instead of hanging."""
    from mmgt_b200.frame_shard import FrameShardGroup
    eng = _engine(dev, torch.bfloat16)
    groups = FrameShardGroup.emulate(eng, 2, 4096)
    try:
        streams = [torch.cuda.Stream() for _ in range(2)]
        for _ in range(5):
            for s in (1, 0):
                with torch.cuda.stream(streams[s]):
                    groups[s].barrier()
        for g in groups:
            g.check()
        groups[0]._barrier.timeout_ms = 50
        groups[0].barrier()                      # shard 1 never signals epoch 6
        with pytest.raises(RuntimeError, match="timed out"):
            groups[0].check()
    finally:
        groups[0].close()


def _shard_inputs(win, B, F, k, s):
    """This shard's frames of a window's inputs (frames are the second axis of every per-frame tensor)."""
    Fl = F // k
    sl = slice(s * Fl, (s + 1) * Fl)
    out = dict(win)
    out["sample"] = win["sample"][:, :, sl].contiguous()
    out["pose_cond_fea"] = win["pose_cond_fea"][:, :, sl].contiguous()
    out["audio_embedding"] = win["audio_embedding"][:, sl].contiguous()
    for name in ("full_mask", "face_mask", "body_mask"):
        out[name] = [m.view(B, F, -1)[:, sl].reshape(B * Fl, -1).contiguous() for m in win[name]]
    return out


def _run_shards(dev, dtype, groups, fn):
    """fn(shard index, engine, group-or-None) on one thread + stream per emulated shard; returns the results."""
    k = len(groups)
    outs, errors = [None] * k, []
    ready = threading.Barrier(k)

    def worker(s):
        try:
            eng = _engine(dev, dtype)
            groups[s].eng = eng
            groups[s]._barrier.timeout_ms = 1500
            with torch.cuda.stream(torch.cuda.Stream()), torch.no_grad():
                fn(s, eng, None)                  # warm this stream's allocator pools / kernels: no cudaMalloc or module
                torch.cuda.current_stream().synchronize()   # load may happen while the other shard waits in a barrier
                ready.wait(timeout=60)
                outs[s] = fn(s, eng, groups[s]).float()
                torch.cuda.current_stream().synchronize()
        except Exception as e:   # noqa: BLE001
            errors.append((s, repr(e)))
    threads = [threading.Thread(target=worker, args=(s,)) for s in range(k)]
    interval = sys.getswitchinterval()
    sys.setswitchinterval(1e-4)       # thousands of short ctypes launches per thread: do not wait 5 ms for the GIL each time
    try:
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=300)
    finally:
        sys.setswitchinterval(interval)
    assert not errors, errors
    for g in groups:
        g.check()
    return outs


@pytest.mark.parametrize("dtype,k,tol", [(torch.float32, 2, 1e-5), (torch.float32, 4, 1e-5), (torch.bfloat16, 2, 1e-5),
                                         (torch.bfloat16, 4, 1e-5)], ids=["f32-k2", "f32-k4", "bf16-k2", "bf16-k4"])
def test_motion_module_frame_sharded_matches_unsharded(dev, dtype, k, tol):
    """One VanillaTemporalModule (live proj_out) on k frame shards vs the unsharded module: the temporal delta
    out - x must agree (bf16: every kernel is row-wise deterministic, so the match is to rounding of the final add)."""
    from mmgt_b200.frame_shard import FrameShardGroup
    from mmgt_b200.motion_module import VanillaTemporalModule
    torch.manual_seed(3)
    C, B, F, H = 128, 2, 8, 8
    mod = VanillaTemporalModule(in_channels=C, num_attention_heads=8, num_transformer_block=1,
                                attention_block_types=("Temporal_Self", "Temporal_Self"), temporal_position_encoding=True,
                                temporal_position_encoding_max_len=32, zero_initialize=False).to(dev)
    eng0 = _engine(dev, dtype)
    x = torch.randn(B, F, H, H, C, generator=torch.Generator().manual_seed(8)).to(dev, dtype)
    ref = mod.run(eng0, x.view(B * F, H, H, C), F).float() - x.view(B * F, H, H, C).float()
    torch.cuda.synchronize()
    groups = FrameShardGroup.emulate(eng0, k, B * (F // k) * H * H * C * 4)
    Fl = F // k
    try:
        def fn(s, eng, group):
            xs = x[:, s * Fl:(s + 1) * Fl].reshape(B * Fl, H, H, C).contiguous()
            return mod.run(eng, xs, Fl, group) - xs
        outs = _run_shards(dev, dtype, groups, fn)
        for s in range(k):
            want = ref.view(B, F, H, H, C)[:, s * Fl:(s + 1) * Fl].reshape(outs[s].shape)
            err = rel_l2(outs[s], want)
            print(f"motion module shard {s}/{k} {dtype}: temporal delta rel-L2 vs unsharded {err:.3e}")
            assert err < (tol if dtype == torch.float32 else 2e-2)     # bf16: the delta is a difference of bf16-rounded sums
    finally:
        groups[0].close()


@pytest.mark.parametrize("compute_dtype,k", [(torch.float32, 2), (torch.bfloat16, 2), (torch.bfloat16, 4)],
                         ids=["f32-k2", "bf16-k2", "bf16-k4"])
def test_tiny_unet_frame_sharded_matches_unsharded(dev, compute_dtype, k):
    """A CFG window (B=2, F=8) run as k frame shards -- every motion module exchanging rows through the fused GEMM
    epilogue (bf16) or GEMM + exchange copy (float32) and the flag barrier -- against the unsharded forward:
    float32 to 2e-5; bf16 (where one differently rounded element re-rounds everything downstream) as close to the
    float32 result as the unsharded bf16 forward is."""
    from mmgt_b200.frame_shard import FrameShardGroup
    spec = UNetSpec(block_out_channels=TINY)
    sd = synthetic_state_dict("tiny")
    B, F, latent = 2, 8, 16
    inp = make_inputs(spec, F, latent)
    win = to_dev(window_inputs(inp, list(range(F))), "cuda")
    t = torch.tensor(500)

    def make(dtype):
        unet = build_cuda_unet(TINY, sd, compute_dtype=dtype)
        unet.train()
        unet.enable_gradient_checkpointing()
        attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
        return unet

    def fwd(unet, w, frames, eng, shard):
        x = eng.ncfhw_to_tokens(w["sample"])
        pose = eng.ncfhw_to_tokens(w["pose_cond_fea"])
        with torch.no_grad():
            return unet.forward_tokens(eng, x, t, w["encoder_hidden_states"], w["audio_embedding"], pose, w["full_mask"],
                                       w["face_mask"], w["body_mask"], w["motion_scale"], B, frames, shard=shard)
    unet = make(compute_dtype)
    base_eng = unet._engine(dev)
    ref = fwd(unet, win, F, base_eng, None).float()                # also builds every weight pack / bank projection
    torch.cuda.synchronize()
    if compute_dtype == torch.bfloat16:
        u32 = make(torch.float32)
        exact = fwd(u32, win, F, u32._engine(dev), None).float()
        base_err = rel_l2(ref, exact)
        del u32
    groups = FrameShardGroup.emulate(base_eng, k, B * (F // k) * latent * latent * TINY[0] * 4)
    Fl = F // k
    try:
        outs = _run_shards(dev, compute_dtype, groups,
                           lambda s, eng, group: fwd(unet, _shard_inputs(win, B, F, k, s), Fl, eng, group))
        got = torch.stack([o.view(B, Fl, latent, latent, -1) for o in outs], dim=1).reshape(ref.shape)
        if compute_dtype == torch.float32:
            err = rel_l2(got, ref)
            print(f"frame shards k={k} float32: rel-L2 vs unsharded {err:.3e}")
            assert err < 2e-5
        else:
            err = rel_l2(got, exact)
            print(f"frame shards k={k} bf16: rel-L2 vs float32 {err:.3e} (unsharded bf16: {base_err:.3e}); "
                  f"vs unsharded bf16 {rel_l2(got, ref):.3e}")
            assert err < 1.5 * base_err + 1e-3
    finally:
        groups[0].close()
    del unet
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_schedules_match_oracle(n):
    """One process per GPU, CUDA IPC + NVLink peer stores: tests/multigpu_frame_shard.py under torchrun on n GPUs (every
    schedule of DenoiseLoop at that world size vs the oracle loop).  Skipped when fewer GPUs are visible."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs (run: gpurun --gpus {n} -- python -m pytest tests/test_frame_shard_gpu.py -m gpu -k multi_gpu)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + (os.getpid() + n) % 300), os.path.join(here, "multigpu_frame_shard.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0
    assert "FRAME-SHARD PARITY OK" in r.stdout
