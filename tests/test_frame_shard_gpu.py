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
        plain = [eng.gemm(src_A[s].contiguous(), W, bias=bias, residual=src_R[s].contiguous()) for s in range(k)]
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


@pytest.mark.parametrize("compute_dtype,k,tol", [(torch.bfloat16, 2, 2e-3), (torch.bfloat16, 4, 2e-3), (torch.float32, 2, 2e-5)],
                         ids=["bf16-k2", "bf16-k4", "f32-k2"])
def test_tiny_unet_frame_sharded_matches_unsharded(dev, compute_dtype, k, tol):
    """A CFG window (B=2, F=8) run as k frame shards -- every motion module exchanging rows through the fused GEMM
    epilogue (bf16) or GEMM + exchange copy (float32) and the flag barrier -- equals the unsharded forward."""
    from mmgt_b200.frame_shard import FrameShardGroup
    spec = UNetSpec(block_out_channels=TINY)
    sd = synthetic_state_dict("tiny")
    unet = build_cuda_unet(TINY, sd, compute_dtype=compute_dtype)
    unet.train()
    unet.enable_gradient_checkpointing()
    B, F, latent = 2, 8, 16
    inp = make_inputs(spec, F, latent)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    win = to_dev(window_inputs(inp, list(range(F))), "cuda")
    t = torch.tensor(500)

    def fwd(w, frames, eng, shard):
        x = eng.ncfhw_to_tokens(w["sample"])
        pose = eng.ncfhw_to_tokens(w["pose_cond_fea"])
        with torch.no_grad():
            return unet.forward_tokens(eng, x, t, w["encoder_hidden_states"], w["audio_embedding"], pose, w["full_mask"],
                                       w["face_mask"], w["body_mask"], w["motion_scale"], B, frames, shard=shard)
    base_eng = unet._engine(dev)
    ref = fwd(win, F, base_eng, None).float()                      # also builds every weight pack / bank projection
    torch.cuda.synchronize()
    groups = FrameShardGroup.emulate(base_eng, k, B * (F // k) * latent * latent * TINY[0] * 4)
    outs, errors = [None] * k, []
    ready = threading.Barrier(k)

    def worker(s):
        try:
            eng = _engine(dev, compute_dtype)
            groups[s].eng = eng
            groups[s]._barrier.timeout_ms = 1500
            with torch.cuda.stream(torch.cuda.Stream()):
                mine = _shard_inputs(win, B, F, k, s)
                fwd(mine, F // k, eng, None)      # warm this stream's allocator pools / kernels: no cudaMalloc or module
                torch.cuda.current_stream().synchronize()   # load may happen while the other shard waits in a barrier
                ready.wait(timeout=60)
                outs[s] = fwd(_shard_inputs(win, B, F, k, s), F // k, eng, groups[s]).float()
                torch.cuda.current_stream().synchronize()
        except Exception as e:   # noqa: BLE001
            errors.append((s, repr(e)))
    try:
        threads = [threading.Thread(target=worker, args=(s,)) for s in range(k)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=300)
        assert not errors, errors
        for g in groups:
            g.check()
        Fl = F // k
        for s in range(k):
            want = ref.view(B, F, latent, latent, -1)[:, s * Fl:(s + 1) * Fl].reshape(outs[s].shape)
            err = rel_l2(outs[s], want)
            print(f"frame shard {s}/{k} {compute_dtype}: rel-L2 vs unsharded {err:.3e}")
            assert err < tol
    finally:
        groups[0].close()
    del unet
    torch.cuda.empty_cache()


def test_two_gpu_denoise_step_frame_sharded():
    """One process per GPU, CUDA IPC + NVLink peer stores: tests/multigpu_frame_shard.py under torchrun."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_frame_shard_gpu.py -m gpu)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(here, "multigpu_frame_shard.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0
    assert "FRAME-SHARD PARITY OK" in r.stdout
