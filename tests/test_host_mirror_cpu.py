"""CPU: the HOST side of the product -- module orchestration, weight packs, GEGLU interleave, fused MM-HAA weights, window
gather / accumulate, CFG + DDIM, Pose2VideoPipeline.__call__, frame-shard row exchange -- executed on tests/fake_engine.py
(float32 PyTorch stand-in for the C-ABI operators) and compared with the oracle / the reference goldens.  The CUDA kernels
themselves are covered by the -m gpu tests; this file keeps the Python mirror honest where there is no GPU."""
import os
import threading

import numpy as np
import pytest
import torch

from fake_engine import FakeEngine, FakeShardGroup
from helpers import GOLD, TINY, attach_banks, bank_pairing_order, build_cuda_unet, rel_l2, synthetic_state_dict
from oracle.sampler import DDIM, denoise_step, uniform_windows
from oracle.synthetic import make_banks, make_inputs, window_inputs
from oracle.unet3d import UNetSpec, unet3d_forward


def _tiny_unet(cfg=True, scripts_branch=True):
    spec = UNetSpec(block_out_channels=TINY)
    sd = synthetic_state_dict("tiny")
    unet = build_cuda_unet(TINY, sd, device="cpu")
    if scripts_branch:
        unet.train()
        unet.enable_gradient_checkpointing()
    else:
        unet.eval()
    return spec, sd, unet


def _forward(unet, eng, win, t, B, frames, shard=None):
    x = eng.ncfhw_to_tokens(win["sample"])
    pose = eng.ncfhw_to_tokens(win["pose_cond_fea"])
    with torch.no_grad():
        y = unet.forward_tokens(eng, x, torch.tensor(t), win["encoder_hidden_states"], win["audio_embedding"], pose,
                                win["full_mask"], win["face_mask"], win["body_mask"], win["motion_scale"], B, frames, shard=shard)
    return y


@pytest.mark.parametrize("fuse_audio,interleaved", [(True, True), (False, False)], ids=["fused-mmhaa", "per-region"])
@pytest.mark.parametrize("branch", ["scripts", "eval"])
def test_host_mirror_matches_reference_golden_on_cpu(fuse_audio, interleaved, branch):
    """UNet3DConditionModel.forward_tokens on the fake engine vs the outputs of the reference's own modules."""
    g = np.load(os.path.join(GOLD, "unet_tiny.npz"))
    latent, frames, t = int(g["latent"]), int(g["frames"]), int(g["timestep"])
    spec, sd, unet = _tiny_unet(scripts_branch=branch == "scripts")
    inp = make_inputs(spec, frames, latent)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    win = window_inputs(inp, list(range(frames)))
    eng = FakeEngine(fuse_audio=fuse_audio, interleaved_geglu=interleaved)
    y = _forward(unet, eng, win, t, 2, frames)
    out = eng.tokens_to_ncfhw(y, 2, frames, torch.float32)
    err = rel_l2(out, torch.from_numpy(g[f"out_{branch}"]))
    assert err < 2e-5, err
    assert (eng.calls.get("audio_attention", 0) > 0) == fuse_audio          # the fused three-region path really ran


def test_denoise_loop_and_pipeline_call_on_cpu(monkeypatch):
    """DenoiseLoop (windows, CFG, overlap average, DDIM) and Pose2VideoPipeline.__call__ through the fake engine vs the
    oracle loop; also DenoiseLoop.reload for a second video."""
    from mmgt_b200.pipeline_pose2vid_long import DenoiseLoop, Pose2VideoPipeline
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    spec, sd, unet = _tiny_unet()
    eng = FakeEngine()
    monkeypatch.setattr(unet, "_engine", lambda device: eng)
    L, latent, n_steps = 16, 8, 3
    banks = make_banks(spec, latent)
    windows = uniform_windows(0, L)

    def unet_fn(sample, t, ehs, aud, pose, full, face, lip, ms):
        with torch.no_grad():
            return unet3d_forward(sd, spec, sample, t, ehs, aud, pose, full, face, lip, ms, banks, ref_index=[None, 1],
                                  apply_motion_scale=True)

    def oracle(inp):
        ddim, lat = DDIM(), inp["latents"].clone()
        for t in ddim.timesteps(n_steps):
            lat, _ = denoise_step(unet_fn, lat, t, n_steps, ddim, 3.5, windows, inp["pose_fea"], inp["audio"],
                                  inp["full_mask"], inp["face_mask"], inp["lip_mask"], inp["encoder_hidden_states"],
                                  inp["motion_scale"])
        return lat
    vids = [make_inputs(spec, L, latent, seed=s) for s in (3, 4)]
    refs = [oracle(v) for v in vids]

    def args(d):
        return (d["latents"], d["pose_fea"], d["audio"], d["full_mask"], d["face_mask"], d["lip_mask"], d["encoder_hidden_states"])
    attach_banks(unet, spec, banks, cfg=True)
    loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 3.5, motion_scale=vids[0]["motion_scale"])
    loop.prepare(*args(vids[0]))
    assert loop.windows == windows
    assert rel_l2(loop.run(), refs[0]) < 2e-5
    loop.reload(*args(vids[1]))
    assert rel_l2(loop.run(), refs[1]) < 2e-5

    inp = vids[0]
    pipe = Pose2VideoPipeline(vae=None, image_encoder=None, reference_unet=None, denoising_unet=unet, pose_guider=None,
                              scheduler=DDIMSchedule.from_config())
    cond = lambda ms: [m[:L] for m in ms]   # noqa: E731
    out = pipe(None, None, inp["audio"][1:2], cond(inp["full_mask"]), cond(inp["face_mask"]), cond(inp["lip_mask"]),
               width=latent * 8, height=latent * 8, video_length=L, num_inference_steps=n_steps, guidance_scale=3.5,
               motion_scale=inp["motion_scale"], output_type="latent", clip_image_embeds=inp["encoder_hidden_states"][1],
               pose_fea=inp["pose_fea"], reference_banks=[banks[p] for p in bank_pairing_order(spec)], latents=inp["latents"])
    assert rel_l2(out.videos, refs[0]) < 2e-5


@pytest.mark.parametrize("k", [2, 4])
def test_frame_sharded_unet_on_cpu_matches_unsharded(k):
    """A CFG window split over k emulated shards (threads): the production host code of the sharded motion modules and the
    production row mapping, rows exchanged through FakeShardGroup mailboxes, vs the unsharded forward."""
    spec, sd, unet = _tiny_unet()
    B, frames, latent = 2, 8, 16
    inp = make_inputs(spec, frames, latent, seed=9)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    win = window_inputs(inp, list(range(frames)))
    ref = _forward(unet, FakeEngine(), win, 500, B, frames)
    Fl = frames // k
    groups = FakeShardGroup.make(k)
    outs, errors = [None] * k, []

    def shard_inputs(s):
        sl = slice(s * Fl, (s + 1) * Fl)
        w = dict(win)
        w["sample"] = win["sample"][:, :, sl]
        w["pose_cond_fea"] = win["pose_cond_fea"][:, :, sl]
        w["audio_embedding"] = win["audio_embedding"][:, sl]
        for name in ("full_mask", "face_mask", "body_mask"):
            w[name] = [m.view(B, frames, -1)[:, sl].reshape(B * Fl, -1) for m in win[name]]
        return w

    def worker(s):
        try:
            outs[s] = _forward(unet, FakeEngine(), shard_inputs(s), 500, B, Fl, shard=groups[s])
        except Exception as e:   # noqa: BLE001
            errors.append((s, repr(e)))
            groups[s]._barrier.abort()
    threads = [threading.Thread(target=worker, args=(s,)) for s in range(k)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=600)
    assert not errors, errors
    got = torch.stack([o.view(B, Fl, latent, latent, -1) for o in outs], dim=1).reshape(ref.shape)
    assert rel_l2(got, ref) < 1e-5


@pytest.mark.parametrize("world,k,remainder,L,latent", [(4, 1, False, 80, 8), (8, 2, True, 80, 16), (2, 2, False, 20, 16)],
                         ids=["4-ranks-whole", "8-ranks-window+shared-forward", "2-ranks-all-sharded"])
def test_multi_rank_schedules_sum_to_the_single_rank_step(monkeypatch, world, k, remainder, L, latent):
    """bench.py's schedules at 4 and 8 GPUs (and the all-sharded mode), rank by rank on the fake engine: the per-rank
    accumulated predictions (what the per-step all-reduce sums) must add up to the single-rank accumulator.  The ranks of a
    frame-shard group run in threads and exchange rows through FakeShardGroup, exactly the host code of the GPU run."""
    from mmgt_b200.pipeline_pose2vid_long import DenoiseLoop
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    spec, sd, unet = _tiny_unet()
    monkeypatch.setattr(unet, "_engine", lambda device: FakeEngine())
    n_steps = 30
    inp = make_inputs(spec, L, latent, seed=21)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    args = (inp["latents"], inp["pose_fea"], inp["audio"], inp["full_mask"], inp["face_mask"], inp["lip_mask"],
            inp["encoder_hidden_states"])

    def make(rank, world_size, group=None):
        loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 3.5, motion_scale=inp["motion_scale"], rank=rank,
                           world_size=world_size, frame_shards=k if world_size > 1 else 1, shard_group=group,
                           shard_remainder=remainder and world_size > 1)
        loop.prepare(*args)
        loop.t_dev.fill_(float(loop.timesteps[0]))
        return loop
    single = make(0, 1)
    single._units_body()
    total = torch.zeros_like(single.noise_acc)
    work = []
    for g0 in range(0, world, k):
        groups = FakeShardGroup.make(k) if k > 1 else [None]
        loops = [make(g0 + s, world, groups[s]) for s in range(k)]
        errors = []

        def body(lp):
            try:
                lp._units_body()
            except Exception as e:   # noqa: BLE001
                errors.append(repr(e))
                if lp.shard_group is not None:
                    lp.shard_group._barrier.abort()
        threads = [threading.Thread(target=body, args=(lp,)) for lp in loops]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=900)
        assert not errors, errors
        for lp in loops:
            total += lp.noise_acc
            work.append(sum(len(b) * e["frames"] / 12.0 for (_, b, _), e in zip(lp.units, lp.prepared)))
    assert rel_l2(total, single.noise_acc) < 1e-5
    if L == 80:
        assert max(work) == min(work) == 20.0 / world       # balanced: 5 (4 ranks) / 2.5 (8 ranks) forwards of work per rank
