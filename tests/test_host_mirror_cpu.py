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


@pytest.mark.parametrize("fuse_audio,interleaved,ln_fused", [(True, True, True), (True, True, False), (False, False, False)],
                         ids=["fused-mmhaa-foldedLN-subpixel", "fused-mmhaa", "per-region"])
@pytest.mark.parametrize("branch", ["scripts", "eval"])
def test_host_mirror_matches_reference_golden_on_cpu(fuse_audio, interleaved, ln_fused, branch):
    """UNet3DConditionModel.forward_tokens on the fake engine vs the outputs of the reference's own modules."""
    g = np.load(os.path.join(GOLD, "unet_tiny.npz"))
    latent, frames, t = int(g["latent"]), int(g["frames"]), int(g["timestep"])
    spec, sd, unet = _tiny_unet(scripts_branch=branch == "scripts")
    inp = make_inputs(spec, frames, latent)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    win = window_inputs(inp, list(range(frames)))
    eng = FakeEngine(fuse_audio=fuse_audio, interleaved_geglu=interleaved, ln_fused=ln_fused)
    y = _forward(unet, eng, win, t, 2, frames)
    out = eng.tokens_to_ncfhw(y, 2, frames, torch.float32)
    err = rel_l2(out, torch.from_numpy(g[f"out_{branch}"]))
    assert err < 2e-5, err
    assert (eng.calls.get("audio_attention", 0) > 0) == fuse_audio          # the fused three-region path really ran
    # LayerNorm folded into the consuming GEMMs (row statistics only) vs LayerNorm passes
    assert (eng.calls.get("row_stats", 0) > 0) == ln_fused and (eng.calls.get("layernorm", 0) == 0) == ln_fused


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
    # the CFG branches of every window as separate single-branch units, levels batched 1 / 2 / all / all units at a time
    loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 3.5, motion_scale=vids[0]["motion_scale"])
    loop.split_branches, loop.level_batch = True, [1, 2, 99, 99]
    loop.prepare(*args(vids[0]))
    assert [len(b) for _, b, _ in loop.units] == [1] * (2 * len(windows))
    assert rel_l2(loop.run(), refs[0]) < 2e-5

    inp = vids[0]
    pipe = Pose2VideoPipeline(vae=None, image_encoder=None, reference_unet=None, denoising_unet=unet, pose_guider=None,
                              scheduler=DDIMSchedule.from_config())
    cond = lambda ms: [m[:L] for m in ms]   # noqa: E731
    out = pipe(None, None, inp["audio"][1:2], cond(inp["full_mask"]), cond(inp["face_mask"]), cond(inp["lip_mask"]),
               width=latent * 8, height=latent * 8, video_length=L, num_inference_steps=n_steps, guidance_scale=3.5,
               motion_scale=inp["motion_scale"], output_type="latent", clip_image_embeds=inp["encoder_hidden_states"][1],
               pose_fea=inp["pose_fea"], reference_banks=[banks[p] for p in bank_pairing_order(spec)], latents=inp["latents"])
    assert rel_l2(out.videos, refs[0]) < 2e-5


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _StubClip(torch.nn.Module):
    """CLIPVisionModelWithProjection stand-in: (1, 3, 224, 224) -> .image_embeds (1, 768)."""

    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Linear(3, 768)

    @property
    def dtype(self):
        return self.proj.weight.dtype

    def forward(self, pixel_values):
        return _Out(image_embeds=self.proj(pixel_values.mean(dim=(2, 3))))


class _StubVae(torch.nn.Module):
    """AutoencoderKL stand-in: 8x average pool to 4 channels; records what it was asked to encode."""

    def __init__(self):
        super().__init__()
        self.mix = torch.nn.Conv2d(3, 4, 1)
        self.config = _Out(block_out_channels=[128, 256, 512, 512])
        self.seen = None

    @property
    def dtype(self):
        return self.mix.weight.dtype

    @property
    def device(self):
        return self.mix.weight.device

    def encode(self, x):
        self.seen = x
        return _Out(latent_dist=_Out(mean=torch.nn.functional.avg_pool2d(self.mix(x), 8)))


class BasicTransformerBlock(torch.nn.Module):
    """Named like the reference's 2-D block (attention.py:12): the write-mode controller finds it by name + norm1."""

    def __init__(self, dim):
        super().__init__()
        self.norm1 = torch.nn.LayerNorm(dim)
        self.lin = torch.nn.Linear(dim, dim)

    def forward(self, hidden_states, encoder_hidden_states=None):
        return hidden_states + self.lin(self.norm1(hidden_states))


class _StubReferenceNet(torch.nn.Module):
    """2-D UNet stand-in with the SD-1.5 transformer-block layout (widths / token counts of the denoising UNet's 16
    spatial blocks, in module order down -> up -> mid like the reference's UNet2DConditionModel)."""

    def __init__(self, layout):
        super().__init__()
        self.layout = layout                               # [(width, tokens)] in module order
        self.inp = torch.nn.ModuleList([torch.nn.Linear(4, c) for c, _ in layout])
        self.blocks = torch.nn.ModuleList([BasicTransformerBlock(c) for c, _ in layout])
        self.calls = []

    @property
    def dtype(self):
        return self.blocks[0].lin.weight.dtype

    def forward(self, latents, timestep, encoder_hidden_states=None, return_dict=True):
        self.calls.append((tuple(latents.shape), float(timestep), tuple(encoder_hidden_states.shape)))
        B = latents.shape[0]
        for (c, t), inp, blk in zip(self.layout, self.inp, self.blocks):
            side = int(round(t ** 0.5))
            x = torch.nn.functional.adaptive_avg_pool2d(latents, side).flatten(2).transpose(1, 2)      # (B, t, 4)
            x = inp(x) + encoder_hidden_states.mean(dim=(1, 2)).view(B, 1, 1)                           # cond rows differ
            blk(x, encoder_hidden_states=encoder_hidden_states)
        return (latents,)


class _StubPoseGuider(torch.nn.Module):
    def __init__(self, width):
        super().__init__()
        self.conv = torch.nn.Conv3d(3, width, 1)
        self.seen = None

    @property
    def dtype(self):
        return self.conv.weight.dtype

    def forward(self, cond):
        self.seen = cond
        B, C, L, H, W = cond.shape
        y = torch.nn.functional.avg_pool3d(cond, (1, 8, 8))
        return 0.1 * self.conv(y)


def test_pipeline_call_with_the_reference_scripts_argument_list_on_cpu(monkeypatch):
    """Pose2VideoPipeline.__call__ exactly as scripts/pose2vid.py:284-296 calls it -- PIL reference image, list of PIL pose
    images, mask lists, zero audio -- with stub CLIP / VAE / ReferenceNet / PoseGuider modules: CLIP embedding, VAE
    encode, ReferenceNet write pass (pre-hook banks), update(), pose features, CFG duplication and the loop, vs the
    oracle loop fed with the same conditioning.  A second call (new reference image) must reuse the cached loop."""
    from PIL import Image
    from mmgt_b200.image_processor import VaeImageProcessor
    from mmgt_b200.mutual_self_attention import _reader_blocks
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    spec, sd, unet = _tiny_unet()
    eng = FakeEngine()
    monkeypatch.setattr(unet, "_engine", lambda device: eng)
    L, latent, n_steps = 16, 8, 2
    size = latent * 8
    torch.manual_seed(5)
    # reference-net layout = the reader blocks in MODULE order (the controllers sort both sides the same stable way)
    readers_module_order = unet.spatial_blocks()
    tok = {TINY[0]: latent ** 2, TINY[1]: (latent // 2) ** 2}
    names = {id(m): n for n, m in unet.named_modules()}
    layout = []
    for b in readers_module_order:
        c = b.norm1.normalized_shape[0]
        n = names[id(b)]
        if c in tok and c != TINY[2]:
            t = tok[c]
        else:   # widest level: 8x8 -> 2x2 at the third level, 1x1 at the mid block
            t = (latent // 8) ** 2 if n.startswith("mid_block") else (latent // 4) ** 2
        layout.append((c, t))
    vae, clip, refnet, guider = _StubVae(), _StubClip(), _StubReferenceNet(layout), _StubPoseGuider(TINY[0])
    pipe = Pose2VideoPipeline(vae=vae, image_encoder=clip, reference_unet=refnet, denoising_unet=unet, pose_guider=guider,
                              scheduler=DDIMSchedule.from_config())
    rng = np.random.default_rng(0)

    def pil(seed_shift=0):
        return Image.fromarray(rng.integers(0, 256, (size + 8 * seed_shift, size, 3), dtype=np.uint8), "RGB")
    ref_image, pose_list = pil(1), [pil() for _ in range(L)]        # the reference image gets resized to (size, size)
    inp = make_inputs(spec, L, latent, seed=3)
    cond = lambda ms: [m[L:] for m in ms]   # noqa: E731  (the cond half = what a script passes before CFG duplication)
    gen = torch.Generator().manual_seed(42)
    zero_audio = torch.zeros(1, L, 32, 768)

    def call(image):
        return pipe(ref_image=image, pose_images=pose_list, audio_tensor=zero_audio,
                    pixel_values_full_mask=cond(inp["full_mask"]), pixel_values_face_mask=cond(inp["face_mask"]),
                    pixel_values_lip_mask=cond(inp["lip_mask"]), width=size, height=size, video_length=L,
                    num_inference_steps=n_steps, guidance_scale=3.5, generator=gen, motion_scale=[1.0, 1.0, 2.0],
                    output_type="latent").videos

    def expected(image, latents0):
        """The same conditioning computed by hand, pushed through the oracle loop."""
        with torch.no_grad():
            from transformers import CLIPImageProcessor
            px = CLIPImageProcessor().preprocess(image.resize((224, 224)), return_tensors="pt").pixel_values
            e = clip(px).image_embeds.unsqueeze(1)
            ehs = torch.cat([torch.zeros_like(e), e])
            ref = VaeImageProcessor(8, do_convert_rgb=True).preprocess(image, height=size, width=size)
            ref_lat = vae.encode(ref).latent_dist.mean * 0.18215
            feats = []
            for (c, t), inp_l, blk in zip(refnet.layout, refnet.inp, refnet.blocks):
                side = int(round(t ** 0.5))
                x = torch.nn.functional.adaptive_avg_pool2d(ref_lat.repeat(2, 1, 1, 1), side).flatten(2).transpose(1, 2)
                feats.append(blk.norm1(inp_l(x) + ehs.mean(dim=(1, 2)).view(2, 1, 1)).half().float())
            # pair like the controllers: both sides stable-sorted by descending width
            order = sorted(range(len(feats)), key=lambda i: -layout[i][0])
            sorted_readers = _reader_blocks(unet, "full")
            banks = {}
            for i, r in zip(order, sorted_readers):
                banks[names[id(r)].replace(".transformer_blocks.0", "")] = feats[i]
            poses = torch.cat([VaeImageProcessor(8, do_convert_rgb=True, do_normalize=False)
                               .preprocess(p, height=size, width=size).unsqueeze(2) for p in pose_list], dim=2)
            pose_fea = guider(poses)
        dup = lambda ms: [torch.cat([m, m]) for m in cond(ms)]   # noqa: E731
        audio = torch.cat([zero_audio, zero_audio])

        def unet_fn(sample, t, ehs_, aud, pose, full, face, lip, ms):
            with torch.no_grad():
                return unet3d_forward(sd, spec, sample, t, ehs_, aud, pose, full, face, lip, ms, banks, ref_index=[None, 1],
                                      apply_motion_scale=True)
        ddim, lat = DDIM(), latents0.clone()
        for t in ddim.timesteps(n_steps):
            lat, _ = denoise_step(unet_fn, lat, t, n_steps, ddim, 3.5, uniform_windows(0, L), pose_fea, audio,
                                  dup(inp["full_mask"]), dup(inp["face_mask"]), dup(inp["lip_mask"]), ehs, [1.0, 1.0, 2.0])
        return lat
    lat0 = torch.randn((1, 4, L, latent, latent), generator=torch.Generator().manual_seed(42))
    out1 = call(ref_image)
    assert tuple(vae.seen.shape) == (1, 3, size, size) and float(vae.seen.min()) < 0        # resized, normalised to [-1, 1]
    assert tuple(guider.seen.shape) == (1, 3, L, size, size) and float(guider.seen.min()) >= 0
    assert refnet.calls == [((2, 4, latent, latent), 0.0, (2, 1, 768))]
    assert all(len(b.bank) == 0 for b in refnet.blocks) and not any(b._forward_pre_hooks for b in refnet.blocks)
    assert rel_l2(out1, expected(ref_image, lat0)) < 2e-5
    # second video of the same shape, another reference image: the cached loop is reloaded (new banks, new CLIP vector)
    ref2 = pil(2)
    lat1 = torch.randn((1, 4, L, latent, latent), generator=gen.manual_seed(7))
    gen.manual_seed(7)
    assert len(pipe._loops) == 1
    loop = next(iter(pipe._loops.values()))
    out2 = call(ref2)
    assert len(pipe._loops) == 1 and next(iter(pipe._loops.values())) is loop
    assert rel_l2(out2, expected(ref2, lat1)) < 2e-5
    with pytest.raises(NotImplementedError):
        pipe(ref_image, pose_list, zero_audio, [], [], [], size, size, L, n_steps, 3.5, eta=0.5)


def test_interpolate_latents_matches_the_reference_recipe():
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    pipe = Pose2VideoPipeline(None, None, None, None, None, None)
    x = torch.randn(1, 4, 5, 3, 3)
    assert pipe.interpolate_latents(x, 1) is x
    y = pipe.interpolate_latents(x, 3)
    assert y.shape[2] == 4 * 3 + 1
    assert torch.equal(y[:, :, 0], x[:, :, 0]) and torch.equal(y[:, :, -1], x[:, :, -1]) and torch.equal(y[:, :, 3], x[:, :, 1])
    assert torch.allclose(y[:, :, 4], (1 - 1 / 3) * x[:, :, 1] + (1 / 3) * x[:, :, 2])


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


def _conditioning_modules():
    import json
    from mmgt_b200.audio_proj import AudioProjModel
    from mmgt_b200.pose_guider import PoseGuider
    from oracle.make_golden_f1 import AUDIO_CFG, POSE_CFG, zero_init_visible
    from oracle.weights import make_state_dict
    with open(os.path.join(GOLD, "conditioning_spec.json")) as f:
        spec = json.load(f)
    pg, ap = PoseGuider(**POSE_CFG), AudioProjModel(**AUDIO_CFG)
    assert [(k, list(v.shape)) for k, v in pg.state_dict().items()] == [(k, s) for k, s in spec["pose_guider"]]
    assert [(k, list(v.shape)) for k, v in ap.state_dict().items()] == [(k, s) for k, s in spec["audio_proj"]]
    assert float(pg.conv_out.weight.abs().max()) == 0.0                      # zero_module, pose_guider.py:38-45
    pg.load_state_dict(zero_init_visible(make_state_dict([(k, tuple(s)) for k, s in spec["pose_guider"]], seed=3), "conv_out",
                                         seed=31), strict=True)
    ap.load_state_dict(make_state_dict([(k, tuple(s)) for k, s in spec["audio_proj"]], seed=4), strict=True)
    return pg, ap


def test_pose_guider_and_audio_proj_host_mirror_on_cpu():
    """f1: PoseGuider / AudioProjModel (reference constructors and state-dict keys) on the fake engine vs the outputs of the
    reference's own classes (tests/golden/conditioning.npz)."""
    from oracle.make_golden_f1 import audio_input, pose_input
    g = np.load(os.path.join(GOLD, "conditioning.npz"))
    pg, ap = _conditioning_modules()
    eng = FakeEngine()
    pg._engine = ap._engine = lambda x: eng
    assert rel_l2(pg(pose_input()), torch.from_numpy(g["pose_out"])) < 1e-6
    assert rel_l2(ap(audio_input()), torch.from_numpy(g["audio_out"])) < 1e-6


def _decode_worker(rank, world, port, q):
    import torch.distributed as dist
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    from vae_stub import VaeStub
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vae = VaeStub()
    pipe = Pose2VideoPipeline(vae, None, None, None, None, None)
    pipe.rank, pipe.world_size = rank, world
    lat = torch.randn(1, 4, 7, 4, 4, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        out = pipe.decode_latents(lat, decode_chunk_size=2)
    q.put((rank, out, list(vae.decode_calls)))
    dist.destroy_process_group()


def test_decode_latents_chunked_and_frame_sharded_matches_the_sequential_reference_recipe():
    """f2: decode_latents in chunks and as frame slices over 2 ranks (gloo) == the reference's one-frame-at-a-time loop
    (pipeline_pose2vid_long.py:112-125)."""
    import torch.multiprocessing as mp
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    from vae_stub import VaeStub
    vae = VaeStub()
    pipe = Pose2VideoPipeline(vae, None, None, None, None, None)
    assert pipe.vae_scale_factor == 8
    lat = torch.randn(1, 4, 7, 4, 4, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = torch.cat([vae.decode(1 / 0.18215 * lat[:, :, f]).sample for f in range(7)])     # the reference's loop
        ref = ((ref.permute(1, 0, 2, 3).unsqueeze(0) / 2 + 0.5).clamp(0, 1)).numpy()
        vae.decode_calls.clear()
        out = pipe.decode_latents(lat, decode_chunk_size=3)
    assert out.shape == (1, 3, 7, 32, 32) and vae.decode_calls == [3, 3, 1]
    assert np.abs(out - ref).max() < 1e-5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_decode_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][2] == [2, 2] and res[1][2] == [2, 1]                      # 4 + 3 frames, two per call
    for _, o, _ in res:
        assert np.abs(o - ref).max() < 1e-5


@pytest.mark.parametrize("deep_from", [1, 2, 3, [1, 2, 3, 3], [1, 1, 2, 3], [2, 2, 3, 3], [3, 3, 3, 3], [1, 2, 2, 2]])
def test_forward_tokens_group_matches_unit_by_unit_on_cpu(deep_from):
    """Deep-level batching (UNet3DConditionModel.forward_tokens_group): three units -- two CFG windows and one single-branch
    forward -- with the levels from down block ``deep_from`` on (default 2: 16x16 / 8x8 at 512x512) run as one batch vs
    forward_tokens unit by unit (fake engine, float32); a list = units per batch at each level (``level_batch``)."""
    spec, sd, unet = _tiny_unet()
    L, latent, frames = 12, 16, 4
    inp = make_inputs(spec, L, latent, seed=5)
    attach_banks(unet, spec, make_banks(spec, latent), cfg=True)
    eng = FakeEngine()
    units, refs = [], []
    for lo, branches in ((0, (0, 1)), (4, (0, 1)), (8, (1,))):
        win = window_inputs(inp, list(range(lo, lo + frames)))
        sel = list(branches)
        rows = torch.cat([torch.arange(frames) + b * frames for b in branches])
        un = dict(x=eng.ncfhw_to_tokens(win["sample"][sel]), pose=eng.ncfhw_to_tokens(win["pose_cond_fea"][sel]),
                  ehs=win["encoder_hidden_states"][sel], audio=win["audio_embedding"][sel],
                  full=[m[rows] for m in win["full_mask"]], face=[m[rows] for m in win["face_mask"]],
                  body=[m[rows] for m in win["body_mask"]], B=len(branches), F=frames,
                  ref=[None if b == 0 else b for b in branches])
        units.append(un)
        with torch.no_grad():
            refs.append(unet.forward_tokens(eng, un["x"], torch.tensor(500), un["ehs"], un["audio"], un["pose"], un["full"],
                                            un["face"], un["body"], inp["motion_scale"], un["B"], frames, ref_index=un["ref"]))
    eng2 = FakeEngine()
    with torch.no_grad():
        kw = dict(level_batch=deep_from) if isinstance(deep_from, list) else dict(deep_from=deep_from)
        outs = unet.forward_tokens_group(eng2, units, torch.tensor(500), inp["motion_scale"], **kw)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and rel_l2(o, r) < 2e-6
    # fewer operator calls: the deep levels ran once for all three units
    assert eng2.calls["gemm"] < eng.calls["gemm"] and eng2.calls["conv3x3"] < eng.calls["conv3x3"]


def _tiny_refnet(device="cpu"):
    import json
    from mmgt_b200.unet_2d_condition import UNet2DConditionModel
    from oracle.reference_loader import SD15_CFG
    from oracle.weights import make_state_dict
    with open(os.path.join(GOLD, "refnet_spec.json")) as f:
        spec = [(k, tuple(s)) for k, s in json.load(f)]
    cfg = dict(SD15_CFG)
    cfg.update(block_out_channels=list(TINY), down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
               up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, mid_block_type="UNetMidBlock2DCrossAttn")
    net = UNet2DConditionModel.from_config(cfg)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == spec        # the SD-1.5 2-D UNet's 686 tensors
    net.load_state_dict(make_state_dict(spec, seed=8), strict=True)
    return net.to(device)


def _check_refnet_banks(net, tol):
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl, _writer_blocks
    from oracle.make_golden_refnet import refnet_inputs
    g = np.load(os.path.join(GOLD, "refnet_tiny.npz"))
    dev = next(net.parameters()).device
    writer = ReferenceAttentionControl(net, do_classifier_free_guidance=True, mode="write", batch_size=1, fusion_blocks="full")
    x, ehs = refnet_inputs()
    out = net(x.to(dev), torch.zeros((), dtype=torch.long, device=dev), encoder_hidden_states=ehs.to(dev), return_dict=False)[0]
    blocks = _writer_blocks(net, "full")
    assert len(blocks) == 16 and all(len(b.bank) == 1 for b in blocks)
    errs = [rel_l2(b.bank[0], torch.from_numpy(g[f"bank{i:02d}"])) for i, b in enumerate(blocks)]
    e_out = rel_l2(out, torch.from_numpy(g["out"]))
    assert max(errs) < tol and e_out < 2 * tol, (max(errs), e_out)
    writer.clear()
    writer.remove()
    assert all(len(b.bank) == 0 and not b.write_bank for b in blocks)
    return max(errs), e_out


def test_reference_net_write_pass_host_mirror_on_cpu(monkeypatch):
    """f1: the ReferenceNet (SD-1.5 2-D UNet = this package's UNet without motion / audio modules, one frame per sample)
    in write mode on the fake engine vs banks produced by the reference's own classes (oracle/make_golden_refnet.py)."""
    net = _tiny_refnet()
    eng = FakeEngine()
    monkeypatch.setattr(net, "_engine", lambda device: eng)
    _check_refnet_banks(net, 2e-5)


def test_reader_update_takes_the_banks_of_the_references_own_writer_controller():
    """``reader.update(writer)`` with the REFERENCE's ``ReferenceAttentionControl(mode="write")``
    (src/models/mutual_self_attention.py, imported unchanged) on the reference's own network: pairing is by duck typing
    (modules named (Temporal)BasicTransformerBlock that own ``bank`` and ``norm1``, stable sort by descending width,
    mutual_self_attention.py:269-341).  The banks that arrive on this package's reader blocks must be the reference writer's,
    in the committed golden's order, rounded to fp16 like mutual_self_attention.py:340.  Needs the reference checkout."""
    from oracle import reference_loader as RL
    if not RL.reference_available():
        pytest.skip("reference checkout not present (GPU box)")
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl, _reader_blocks
    from oracle.make_golden_refnet import refnet_inputs
    from oracle.weights import make_state_dict
    mods = RL.load_reference_modules()
    cfg = dict(RL.SD15_CFG)
    cfg["block_out_channels"] = list(TINY)
    extra = dict(RL.UNET_ADDITIONAL_KWARGS)
    extra.update(use_motion_module=False, use_audio_module=False)
    ref_net = mods["unet_3d"].UNet3DConditionModel.from_config(cfg, **extra)
    ref_net.load_state_dict(make_state_dict([(k, tuple(v.shape)) for k, v in ref_net.state_dict().items()], seed=8), strict=True)
    ref_net.eval()
    ref_writer = mods["mutual_self_attention"].ReferenceAttentionControl(
        ref_net, do_classifier_free_guidance=True, mode="write", batch_size=1, fusion_blocks="full")
    x, ehs = refnet_inputs()
    with torch.no_grad():
        ref_net(x.unsqueeze(2), torch.zeros((), dtype=torch.long), encoder_hidden_states=ehs, return_dict=False)

    _, _, unet = _tiny_unet()
    reader = ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", batch_size=1, fusion_blocks="full")
    reader.update(ref_writer)                                   # the reference's controller object, not ours
    g = np.load(os.path.join(GOLD, "refnet_tiny.npz"))
    blocks = _reader_blocks(unet, "full")
    assert len(blocks) == 16
    for i, b in enumerate(blocks):
        assert len(b.bank) == 1 and b.bank[0].dtype == torch.float16
        gold = torch.from_numpy(g[f"bank{i:02d}"])
        assert b.bank[0].shape == gold.shape
        assert rel_l2(b.bank[0].float(), gold) < 1e-3, i        # fp16 rounding of the bank: 2^-11 relative
    reader.clear()
    assert all(len(b.bank) == 0 for b in blocks)
    # a writer whose forward has not run yet is refused instead of pairing empty banks
    ref_writer.clear()
    with pytest.raises(ValueError):
        reader.update(ref_writer)
