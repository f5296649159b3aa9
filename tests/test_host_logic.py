"""CPU: host-side mirror of the reference interface, C-ABI surface, packing, multi-rank partitioning."""
import ctypes
import inspect
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import FULL, GOLD, TINY, statedict_spec
from oracle.reference_loader import SD15_CFG, UNET_ADDITIONAL_KWARGS
from oracle.sampler import DDIM, uniform_windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mmgt_b200.h")).read()
    return sorted(set(re.findall(r"MMGT_API\s+[\w\s\*]+?\b(mmgt_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from mmgt_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(lib_built)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mmgt_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"
    assert _lib.load_library().mmgt_abi_version() == 3


def test_library_is_sm100a_tensor_core_code(lib_built):
    out = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out or "sm_100" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, f"{mnemonic} missing: the tcgen05/TMA path was not compiled"


def test_no_gpu_means_loud_failure():
    from mmgt_b200.kernels import get_engine
    with pytest.raises(RuntimeError):
        get_engine(torch.device("cpu"), torch.float32)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmgt_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


@pytest.mark.parametrize("tag,boc", [("tiny", TINY), ("full", FULL)])
def test_state_dict_matches_reference(tag, boc):
    from mmgt_b200.unet_3d import UNet3DConditionModel
    cfg = dict(SD15_CFG)
    cfg["block_out_channels"] = list(boc)
    unet = UNet3DConditionModel.from_config(cfg, **UNET_ADDITIONAL_KWARGS)
    mine = [(k, tuple(v.shape)) for k, v in unet.state_dict().items()]
    assert mine == statedict_spec(tag)
    if tag == "full":
        assert sum(p.numel() for p in unet.parameters()) == 1404718404
    assert unet.training and unet.in_channels == 4 and unet.config.center_input_sample is False
    unet.enable_gradient_checkpointing()
    assert unet._motion_scale_reaches_audio()
    unet.eval()
    assert not unet._motion_scale_reaches_audio()


def test_forward_signature_matches_reference():
    from mmgt_b200.unet_3d import UNet3DConditionModel
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    names = list(inspect.signature(UNet3DConditionModel.forward).parameters)
    assert names[:16] == ["self", "sample", "timestep", "encoder_hidden_states", "audio_embedding", "class_labels",
                          "mask_cond_fea", "pose_cond_fea", "attention_mask", "full_mask", "face_mask", "body_mask",
                          "motion_scale", "down_block_additional_residuals", "mid_block_additional_residual", "return_dict"]
    call = list(inspect.signature(Pose2VideoPipeline.__call__).parameters)
    assert call[:12] == ["self", "ref_image", "pose_images", "audio_tensor", "pixel_values_full_mask",
                         "pixel_values_face_mask", "pixel_values_lip_mask", "width", "height", "video_length",
                         "num_inference_steps", "guidance_scale"]
    p = inspect.signature(Pose2VideoPipeline.__call__).parameters
    assert p["context_frames"].default == 12 and p["context_overlap"].default == 4 and p["context_stride"].default == 1


def test_reference_control_pairing_order():
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl, _reader_blocks
    from mmgt_b200.unet_3d import UNet3DConditionModel
    from oracle.unet3d import UNetSpec, bank_pairing_order
    cfg = dict(SD15_CFG)
    cfg["block_out_channels"] = list(TINY)
    unet = UNet3DConditionModel.from_config(cfg, **UNET_ADDITIONAL_KWARGS)
    ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
    names = {id(m): n for n, m in unet.named_modules()}
    got = [names[id(b)].replace(".transformer_blocks.0", "") for b in _reader_blocks(unet, "full")]
    assert got == bank_pairing_order(UNetSpec(block_out_channels=TINY))


def test_context_and_schedule_mirror():
    from mmgt_b200.context import uniform
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    with open(os.path.join(GOLD, "windows.json")) as f:
        gold = json.load(f)
    for L, w in gold.items():
        assert [list(c) for c in uniform(0, 30, int(L), 12, 1, 4)] == w
    s, d = DDIMSchedule.from_config(), DDIM()
    assert s.timesteps(30) == d.timesteps(30)
    x, v = torch.randn(2, 3), torch.randn(2, 3)
    for t in s.timesteps(30):
        cx, cv = s.step_coefficients(t, 30)
        assert torch.allclose(cx * x + cv * v, d.step(v, t, x, 30), atol=1e-6)


def test_geglu_interleave_roundtrip():
    from mmgt_b200.packing import geglu_interleave
    torch.manual_seed(0)
    n, k, gb = 64, 8, 16
    w, b = torch.randn(2 * n, k), torch.randn(2 * n)
    wi, bi = geglu_interleave(w, b, gb)
    x = torch.randn(5, k)
    ref = torch.nn.functional.linear(x, w, b)
    ref = ref[:, :n] * torch.nn.functional.gelu(ref[:, n:])
    y = torch.nn.functional.linear(x, wi, bi).view(5, n // gb, 2, gb)
    out = (y[:, :, 0] * torch.nn.functional.gelu(y[:, :, 1])).reshape(5, n)
    assert torch.allclose(out, ref, atol=1e-5)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, nb = 80, 2
    windows = uniform_windows(0, L)
    units = [(wi, b) for wi in range(len(windows)) for b in range(nb)][rank::world]
    acc = torch.zeros(nb, 4, L, 2, 2)
    for wi, b in units:   # stand-in prediction that depends on (window, branch, frame)
        for j, f in enumerate(windows[wi]):
            acc[b, :, f] += (wi + 1) * 0.5 + b * 10 + j
    dist.all_reduce(acc)
    q.put((rank, len(units), acc))
    dist.destroy_process_group()


def test_unit_partition_and_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    L, nb = 80, 2
    windows = uniform_windows(0, L)
    ref = torch.zeros(nb, 4, L, 2, 2)
    for wi, c in enumerate(windows):
        for b in range(nb):
            for j, f in enumerate(c):
                ref[b, :, f] += (wi + 1) * 0.5 + b * 10 + j
    assert sorted(r[1] for r in res) == [10, 10]
    for _, _, acc in res:
        assert torch.allclose(acc, ref)


# ---------------------------------------------------------------------------------- frame shards (SURVEY 8e level 3)
def _forwards(groups):
    return sorted((wi, b) for g in groups for wi, bs in g for b in bs)


def test_plan_units_balances_groups():
    """Whole B=2 windows as far as they go, then single-branch forwards, every (window, branch) exactly once."""
    from mmgt_b200.pipeline_pose2vid_long import plan_units
    every = [(wi, b) for wi in range(10) for b in range(2)]
    assert plan_units(10, 2, 1) == [[(wi, (0, 1)) for wi in range(10)]]
    assert [[len(b) for _, b in g] for g in plan_units(10, 2, 2)] == [[2] * 5] * 2
    assert [[len(b) for _, b in g] for g in plan_units(10, 2, 4)] == [[2, 2, 1]] * 4        # 2 windows + 1 forward per rank
    eight = plan_units(10, 2, 8)                                                              # 1 window each + 4 left over
    assert sorted(sum(len(b) for _, b in g) for g in eight) == [2, 2, 2, 2, 3, 3, 3, 3]
    for n in (1, 2, 3, 4, 5, 6, 8):
        assert _forwards(plan_units(10, 2, n)) == every
    assert _forwards(plan_units(3, 1, 2)) == [(0, 0), (1, 0), (2, 0)]                         # no CFG: windows are the forwards


@pytest.mark.parametrize("k,B,F,T", [(2, 1, 12, 16), (4, 2, 12, 64), (2, 2, 4, 6), (3, 1, 6, 9)])
def test_exchange_mapping_is_the_temporal_rearrange(k, B, F, T):
    """direction 1 == '(b f) t c -> all frames of a pixel chunk', direction 2 is its inverse: together they are the
    (b f) d c <-> (b d) f c rearranges of motion_module.py:361-363,386 split over k shards."""
    from mmgt_b200.frame_shard import exchange_destination
    Fl, Tc, C = F // k, T // k, 3
    x = torch.arange(B * F * T * C, dtype=torch.float32).view(B, F, T, C)
    frame_shards = [x[:, s * Fl:(s + 1) * Fl].reshape(-1, C) for s in range(k)]          # rows (b, f_loc, t)
    recv = [torch.full((B * F * Tc, C), -1.0) for _ in range(k)]
    for my in range(k):
        for m in range(B * Fl * T):
            s, row = exchange_destination(1, m, k, my, B, F, T)
            recv[s][row] = frame_shards[my][m]
    for s in range(k):
        assert torch.equal(recv[s], x[:, :, s * Tc:(s + 1) * Tc].reshape(-1, C))         # rows (b, f, t_loc)
    back = [torch.full((B * Fl * T, C), -1.0) for _ in range(k)]
    for my in range(k):
        for m in range(B * F * Tc):
            s, row = exchange_destination(2, m, k, my, B, F, T)
            back[s][row] = recv[my][m]
    for s in range(k):
        assert torch.equal(back[s], frame_shards[s])


def _frame_shard_worker(rank, world, port, q):
    """Two processes run a stand-in 'motion module' (mean over frames per pixel) frame-sharded: rows are exchanged with
    the production row mapping (all_to_all here instead of NVLink peer stores), mixed over frames token-sharded and
    exchanged back; window predictions are overlap-accumulated and all-reduced as DenoiseLoop does."""
    import torch.distributed as dist
    from mmgt_b200.frame_shard import exchange_destination
    from mmgt_b200.pipeline_pose2vid_long import plan_units
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, L, T, C = world, 20, 8, 2
    windows = uniform_windows(0, L)
    video = torch.arange(L * T * C, dtype=torch.float32).view(L, T, C) * 0.01
    units = plan_units(len(windows), 2, world // k)[rank // k]
    acc = torch.zeros(2, L, T, C)

    def exchange(direction, src, B, F):
        Fl, Tc = F // k, T // k
        send = [[] for _ in range(k)]
        for m in range(src.shape[0]):
            s, row = exchange_destination(direction, m, k, rank % k, B, F, T)
            send[s].append((row, src[m]))
        out = torch.zeros_like(src)
        parcels = [None] * k       # gloo has no all_to_all: gather every rank's per-destination parcels, keep ours
        dist.all_gather_object(parcels, [[(r, v.tolist()) for r, v in send[s]] for s in range(k)])
        for src_rank in range(k):
            for r, v in parcels[src_rank][rank % k]:
                out[r] = torch.tensor(v)
        return out

    for wi, branches in units:
        c = windows[wi]
        B, F = len(branches), len(c)
        Fl = F // k
        mine = c[(rank % k) * Fl:(rank % k + 1) * Fl]
        x = torch.stack([video[mine] * (1 + b) for b in branches]).reshape(-1, C)          # rows (b, f_loc, t)
        tok = exchange(1, x, B, F).view(B, F, T // k, C)                                   # rows (b, f, t_loc)
        tok = tok + tok.mean(dim=1, keepdim=True)                                          # mixes ALL frames of a pixel
        y = exchange(2, tok.reshape(-1, C), B, F).view(B, Fl, T, C)
        for bi, b in enumerate(branches):
            for j, f in enumerate(mine):
                acc[b, f] += y[bi, j]
    dist.all_reduce(acc)
    q.put((rank, acc))
    dist.destroy_process_group()


def test_frame_sharded_window_matches_unsharded_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_frame_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    L, T, C = 20, 8, 2
    windows = uniform_windows(0, L)
    video = torch.arange(L * T * C, dtype=torch.float32).view(L, T, C) * 0.01
    ref = torch.zeros(2, L, T, C)
    for c in windows:
        for b in range(2):
            x = video[c] * (1 + b)
            y = x + x.mean(dim=0, keepdim=True)
            for j, f in enumerate(c):
                ref[b, f] += y[j]
    for _, acc in res:
        assert torch.allclose(acc, ref, atol=1e-5)


def test_plan_units_mixed_shares_the_remainder():
    from mmgt_b200.pipeline_pose2vid_long import plan_units_mixed
    every = [(wi, b) for wi in range(10) for b in range(2)]
    whole, shared = plan_units_mixed(10, 2, 8, 2)          # 20 forwards on 8 GPUs: one B=2 window each + 1 forward per pair
    assert [[len(b) for _, b in w] for w in whole] == [[2]] * 8 and [len(s) for s in shared] == [1] * 4
    assert _forwards(whole + shared) == every
    whole, shared = plan_units_mixed(10, 2, 4, 2)          # divides: 2 windows + 1 forward per rank, nothing shared
    assert [[len(b) for _, b in w] for w in whole] == [[2, 2, 1]] * 4 and not any(shared)
    whole, shared = plan_units_mixed(3, 1, 2, 2)           # 3 forwards on 2 GPUs: 1 whole each + 1 shared
    assert [len(w) for w in whole] == [1, 1] and shared == [[(2, (0,))]]
    whole, shared = plan_units_mixed(10, 2, 6, 2)          # 8 single forwards for 6 ranks: 2 left for 3 pairs -> dealt whole
    assert not any(shared) and _forwards(whole) == every
    assert sorted(sum(len(b) for _, b in w) for w in whole) == [3, 3, 3, 3, 4, 4]


@pytest.mark.parametrize("world,k,remainder", [(1, 1, False), (2, 1, False), (4, 1, False), (8, 1, False), (8, 2, True),
                                               (4, 2, True), (2, 2, True), (2, 2, False), (8, 4, False), (6, 2, True)])
def test_every_frame_of_every_forward_is_computed_exactly_once(world, k, remainder):
    """The schedules DenoiseLoop runs (bench.py at 1 / 2 / 4 / 8 GPUs and the explicit frame-shard modes): over all ranks,
    each (window, CFG branch, frame) appears exactly once, and all ranks agree on whether peer buffers are needed."""
    from mmgt_b200.pipeline_pose2vid_long import plan_rank
    windows = uniform_windows(0, 80)
    seen, need, load = [], set(), []
    for rank in range(world):
        units, need_group = plan_rank(len(windows), 2, rank, world, k, remainder)
        need.add(need_group)
        work = 0.0
        for wi, branches, sharded in units:
            c = windows[wi]
            if sharded:
                fl = len(c) // k
                c = c[(rank % k) * fl:(rank % k + 1) * fl]
            seen += [(wi, b, f) for b in branches for f in c]
            work += len(branches) * len(c) / 12.0
        load.append(work)
    assert sorted(seen) == sorted((wi, b, f) for wi, c in enumerate(windows) for b in range(2) for f in c)
    assert len(need) == 1
    if (world, k, remainder) in ((8, 2, True), (4, 2, True), (4, 1, False), (2, 1, False)):
        assert max(load) == min(load) == 20.0 / world          # perfectly balanced: 2.5 forwards per rank on 8 GPUs


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """The ctypes mirrors in mmgt_b200/_lib.py must have the size and field offsets gcc gives the structs of
    include/mmgt_b200.h (a silent mismatch would hand kernels garbage pointers)."""
    import ctypes
    import shutil
    import subprocess
    from mmgt_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"mmgt_gemm_params": _lib.GemmParams, "mmgt_conv3x3_params": _lib.Conv3x3Params,
             "mmgt_attention_params": _lib.AttentionParams, "mmgt_audio_attention_params": _lib.AudioAttentionParams,
             "mmgt_row_exchange": _lib.RowExchange, "mmgt_peer_barrier_params": _lib.PeerBarrierParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mmgt_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_subpixel_upsample_weights_equal_upsample_then_conv():
    """packing.subpixel_upsample_weights: four 2x2-tap convolutions on the low-resolution input == conv3x3(nearest x2)."""
    import torch.nn.functional as F
    from mmgt_b200.packing import subpixel_upsample_weights
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    out = torch.zeros_like(ref)
    H, W = x.shape[2:]
    for (a, b), (taps, dys, dxs) in subpixel_upsample_weights(w).items():
        acc = torch.zeros(2, 4, H, W, dtype=torch.float64)
        xp = F.pad(x, (1, 1, 1, 1))
        for i, dy in enumerate(dys):
            for j, dx in enumerate(dxs):
                patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                acc += torch.einsum("nchw,oc->nohw", patch, taps[:, :, i, j])
        out[:, :, a::2, b::2] = acc
    assert torch.allclose(out, ref, atol=1e-10)


def test_config5_schedule_is_balanced_on_eight_gpus():
    """BASELINE config 5: 160 frames = 20 windows = 40 forwards; 8 GPUs take 2 whole windows + 1 single-branch forward
    each -- balanced without frame shards."""
    from mmgt_b200.pipeline_pose2vid_long import plan_rank
    windows = uniform_windows(0, 160)
    assert len(windows) == 20
    for rank in range(8):
        units, need_group = plan_rank(len(windows), 2, rank, 8, 2, True)
        assert not need_group and [len(b) for _, b, _ in units] == [2, 2, 1] and not any(sh for _, _, sh in units)


def test_engine_wrappers_with_a_mocked_library():
    """Host side of mmgt_b200.kernels.Engine without a GPU: the ctypes parameter blocks it hands to the C ABI and the work
    accounting bench.py reports (the library itself is mocked -- compute is covered by the -m gpu tests)."""
    from unittest import mock
    import mmgt_b200.kernels as K
    eng = object.__new__(K.Engine)
    eng.device, eng.dtype, eng.dt, eng.h, eng.prof, eng.unfused_exchange = torch.device("cpu"), torch.bfloat16, 1, None, None, False
    eng.lib = mock.MagicMock()
    for name in ("mmgt_attention", "mmgt_audio_attention", "mmgt_gemm"):
        getattr(eng.lib, name).return_value = 0
    seen = {}
    eng._t1 = lambda ev, key, flops=0.0, nbytes=0.0: seen.update(key=key, flops=flops, nbytes=nbytes)
    with mock.patch.object(K, "_stream", lambda: None):
        N, T, C, heads = 4, 64, 64, 8
        qkv = torch.zeros(N, T, 3 * C, dtype=torch.bfloat16)
        kv2 = torch.zeros(2, T, 2 * C, dtype=torch.bfloat16)
        idx = torch.tensor([-1, -1, 1, 1], dtype=torch.int32)
        idx._n_seg2 = 2                                       # what UNet3DConditionModel._seg2_index records
        eng.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, k2=kv2[:, :, :C], v2=kv2[:, :, C:], seg2_index=idx)
        p = eng.lib.mmgt_attention.call_args[0][1]._obj
        assert (p.N, p.Lq, p.Lk, p.Lk2, p.heads, p.d, p.B2) == (N, T, T, T, heads, C // heads, 2)
        assert (p.ldq, p.ldk, p.ldk2, p.ldo) == (3 * C, 3 * C, 2 * C, C)
        assert seen["flops"] == 4.0 * heads * T * (N * T + 2 * T) * (C // heads)      # uncond frames: one key segment
        # fused MM-HAA attention: output carries 3C gated columns + 8 gate / pad columns
        q3, kv6 = torch.zeros(N * T, 3 * C, dtype=torch.bfloat16), torch.zeros(N * 32, 6 * C, dtype=torch.bfloat16)
        masks = [torch.ones(N * T) for _ in range(3)]
        out = eng.audio_attention(q3, kv6, masks, (1.0, 1.0, 2.0), N, T, heads)
        a = eng.lib.mmgt_audio_attention.call_args[0][1]._obj
        assert out.shape == (N * T, 3 * C + 8) and (a.N, a.T, a.M, a.heads, a.d) == (N, T, 32, heads, C // heads)
        assert (a.ldq, a.ldkv, a.ldo) == (3 * C, 6 * C, 3 * C + 8) and list(a.scale) == [1.0, 1.0, 2.0]
        # GEMM: strides / GEGLU output width
        A, W = torch.zeros(10, 32, dtype=torch.bfloat16), torch.zeros(64, 32, dtype=torch.bfloat16)
        y = eng.gemm(A, W, geglu_block=16)
        g = eng.lib.mmgt_gemm.call_args[0][1]._obj
        assert y.shape == (10, 32) and (g.M, g.N, g.K, g.lda, g.ldw, g.ldd, g.geglu_block) == (10, 64, 32, 32, 32, 32, 16)
        assert not g.exchange


def test_mask_frontend_matches_the_scripts_recipe():
    """mmgt_b200.mask_frontend (f3): blur_mask + pyramid + the two full-mask recipes against the scripts' own code path --
    cv2 blur (scripts/pose2vid.py:94-114), torchvision Resize + ToTensor per level (image_processor.py:75-102,311-333),
    ``1 + lips`` (scripts/audio2vid.py:470-476) and the per-level ``clamp(1 - face + lips + hands)`` that the broken loop at
    scripts/pose2vid.py:262-271 means.  The pyramid here is the oracle's (numpy); the CUDA one is checked bit-exactly
    against it in tests/test_kernels_gpu.py::test_mask_pyramid_bit_exact."""
    import cv2
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from mmgt_b200 import mask_frontend as mf
    from oracle.mask_pyramid import mask_pyramid_u8

    class OraclePyramid:
        def __init__(self, image_size):
            self.image_size = image_size

        def levels(self, images):
            u8 = np.stack([np.asarray(m, dtype=np.uint8) for m in images])
            return [torch.from_numpy(l.astype(np.float32) / 255.0).reshape(len(images), -1)
                    for l in mask_pyramid_u8(u8, self.image_size)]
    rng = np.random.default_rng(0)
    L = 5

    def frames():
        out = []
        for _ in range(L):
            img = np.zeros((96, 128, 3), dtype=np.uint8)
            y, x = rng.integers(5, 60, 2)
            img[y:y + 30, x:x + 40] = 255
            out.append(Image.fromarray(img))
        return out
    face_f, lips_f, hands_f = frames(), frames(), frames()
    for size in (512, 768):
        full_a, face, lips = mf.motion_masks(face_f, lips_f, None, image_size=size, recipe="audio2vid", pyramid=OraclePyramid(size))
        full_p, _, _ = mf.motion_masks(face_f, lips_f, hands_f, image_size=size, recipe="pose2vid", pyramid=OraclePyramid(size))
        # the scripts' path: cv2 front-end, then torchvision per level
        def script_levels(fr, k):
            pil = []
            for img in fr:
                a = cv2.normalize(cv2.GaussianBlur(cv2.resize(np.array(img), (64, 64)), k, 0), None, 0, 255, cv2.NORM_MINMAX)
                pil.append(Image.fromarray(a.astype(np.uint8)).convert("L"))
            lv = []
            for kk in range(4):
                s = size // (8 << kk)
                t = transforms.Compose([transforms.Resize((s, s)), transforms.ToTensor()])
                lv.append(torch.stack([t(p) for p in pil]).view(L, 1, -1).squeeze(1))
            return lv
        ref_face, ref_lips, ref_hands = script_levels(face_f, (31, 31)), script_levels(lips_f, (21, 21)), script_levels(hands_f, (21, 21))
        for k in range(4):
            assert face[k].shape == (L, (size // (8 << k)) ** 2)
            assert torch.equal(face[k], ref_face[k]) and torch.equal(lips[k], ref_lips[k])          # bit-exact
            assert torch.equal(full_a[k], 1.0 + ref_lips[k])
            assert torch.equal(full_p[k], torch.clamp(1.0 - ref_face[k] + ref_lips[k] + ref_hands[k], 0.0, 1.0))
    with pytest.raises(ValueError):
        mf.full_mask_levels(face, lips, None, "nope")
