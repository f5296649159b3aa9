"""Host mirror of src/pipelines/pipeline_pose2vid_long.py.

``DenoiseLoop`` is the hot loop (:491-646): per DDIM step, one UNet forward per 12-frame context window with
CFG, overlap-averaged, CFG-combined and DDIM-updated -- everything on the sm_100a kernels.
``Pose2VideoPipeline`` keeps the reference's constructor and ``__call__`` signature and argument types; the one-shot
conditioning passes (CLIP, VAE encode, ReferenceNet write pass, pose guider) run on the PyTorch modules handed to the
constructor, exactly where the reference runs them.

Multi-GPU (one process per GPU): the (window, CFG-branch) forwards of a step are independent
 and are dealt round-robin to the rank groups; the only exchange between groups is one
all-reduce of the overlap-accumulated prediction (2*4*L*h*w float32) per step.  With ``frame_shards = k > 1`` a
group is k ranks that split the frames of each window (frame_shard.py): their motion modules switch between the
frame-sharded and the token-sharded layout by storing rows into each other's memory over NVLink.
"""
import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Union

import torch

from .context import get_context_scheduler
from .image_processor import VaeImageProcessor
from .kernels import Engine
from .mutual_self_attention import ReferenceAttentionControl
from .scheduling_ddim import DDIMSchedule


@dataclass
class Pose2VideoPipelineOutput:
    videos: Union[torch.Tensor, "object"]


class DenoiseLoop:
    def __init__(self, unet, schedule: DDIMSchedule, num_inference_steps: int, guidance_scale: float,
                 context_frames: int = 12, context_stride: int = 1, context_overlap: int = 4,
                 context_schedule: str = "uniform", motion_scale: Optional[Sequence[float]] = None,
                 rank: int = 0, world_size: int = 1, process_group=None, frame_shards: int = 1, shard_group=None,
                 shard_remainder: bool = False):
        self.unet = unet
        self.schedule = schedule
        self.n_steps = num_inference_steps
        self.guidance = float(guidance_scale)
        self.cfg = guidance_scale > 1.0
        self.ctx_args = (context_frames, context_stride, context_overlap)
        self.context_scheduler = get_context_scheduler(context_schedule)
        self.motion_scale = motion_scale
        self.rank, self.world, self.group = rank, world_size, process_group
        if frame_shards < 1 or world_size % frame_shards:
            raise ValueError(f"frame_shards={frame_shards} must divide the world size {world_size}")
        self.frame_shards = frame_shards
        # shard_remainder: deal whole (window, branch) forwards to ALL ranks and frame-shard only the ones left over when
        # the count does not divide (20 forwards on 8 GPUs: 2 whole ones per rank + 1 shared by each pair of ranks)
        self.shard_remainder = bool(shard_remainder) and frame_shards > 1
        self.shard_group = shard_group   # frame_shard.FrameShardGroup; created in prepare() unless one is passed in
        self.timesteps = schedule.timesteps(num_inference_steps)
        self._graph = None
        self.graph_launches = 0
        self._made_group = False
        self._bank_ptrs = None
        # units whose 16x16 / 8x8 levels run as ONE batch (UNet3DConditionModel.forward_tokens_group); 1 = unit by unit
        self.deep_batch = 16
        # first down block whose level runs batched over those units (2: the 16x16 / 8x8 levels; 1 adds the 32x32 level:
        # 435.0 vs 440.5 ms per DDIM step, profiles/r2_ab_flags.md run 21 -- the level-1 GEMMs of ONE window are 768 tiles,
        # 5.2 waves of 148 CTAs)
        self.deep_from = 1
        # units per batch at every level, e.g. [1, 2, 99, 99] (overrides deep_from; forward_tokens_group's level_batch)
        self.level_batch = None
        # deal the CFG branches of a whole window as two single-branch units (the 64x64 level of one branch is 31 MB per
        # tensor -- it stays in the 126 MB L2 from the kernel that writes it to the kernel that reads it)
        self.split_branches = False

    # ------------------------------------------------------------------ one-off preparation per video
    def prepare(self, latents, pose_fea, audio, full_mask, face_mask, lip_mask, encoder_hidden_states):
        """latents (1,4,L,h,w); pose_fea (1,320,L,h,w); audio (nb,L,M,768); *_mask: 4 x (nb*L, T_l);
        encoder_hidden_states (nb,1,768) with nb = 2 under CFG ([uncond; cond]) else 1."""
        u = self.unet
        dev = latents.device
        eng: Engine = u._engine(dev)
        self.eng = eng
        nb = 2 if self.cfg else 1
        _, C, L, h, w = latents.shape
        self.L, self.C, self.h, self.w, self.nb = L, C, h, w, nb
        self.latents = latents.to(torch.float32).contiguous().clone()
        cf, cs, co = self.ctx_args
        self.windows = [list(map(int, c)) for c in self.context_scheduler(0, self.n_steps, L, cf, cs, co)]
        # Work units of this rank (plan_rank): whole B=2 windows first, then single-branch forwards, optionally frame shards.
        nw = len(self.windows)
        k = self.frame_shards
        shard = self.rank % k
        self.units, need_group = plan_rank(nw, nb, self.rank, self.world, k, self.shard_remainder)
        if self.split_branches:                # whole (not frame-sharded) windows -> one unit per CFG branch, kept adjacent
            self.units = [v for wi, br, sh in self.units
                          for v in ([(wi, br, sh)] if sh or len(br) == 1 else [(wi, (b,), sh) for b in br])]
        if need_group:
            bad = [len(c) for c in self.windows if len(c) % k]
            if bad:
                raise ValueError(f"frame_shards={k} needs every context window to hold a multiple of {k} frames, got {bad}")
            self._made_group = self.shard_group is None
            if self.shard_group is None:
                from .frame_shard import FrameShardGroup, max_exchange_bytes
                esize = torch.empty((), dtype=eng.dtype).element_size()
                nbr_max = max([len(b) for _, b, sh in self.units if sh] or [1])
                width0 = u.config.block_out_channels[0] if hasattr(u.config, "block_out_channels") else 320
                need = max_exchange_bytes(nbr_max, max(len(c) for c in self.windows) // k, h * w, width0, esize)
                self.shard_group = FrameShardGroup.create(eng, self.rank, self.world, k, need, self.group)
        counts = torch.zeros(L, dtype=torch.float32)
        for c in self.windows:
            for f in c:
                counts[f] += 1
        self.inv_count = (1.0 / counts).to(dev)
        self.noise_acc = torch.zeros((nb, C, L, h, w), device=dev, dtype=torch.float32)
        self.t_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self.prepared = []
        for wi, branches, sharded in self.units:
            c = self.windows[wi]
            if sharded:                   # this rank's frames of the window
                fl = len(c) // k
                c = c[shard * fl:(shard + 1) * fl]
            idx = torch.tensor(c, dtype=torch.int32, device=dev)
            nbr = len(branches)
            x_idx = idx.repeat(nbr)                                                      # latents / pose rows (frames)
            rows = torch.cat([idx + b * L for b in branches])                            # rows of the (nb*L, ...) tensors
            ref = [None if (self.cfg and b == 0) else b for b in branches]
            self.prepared.append(dict(idx=idx, x_idx=x_idx, rows=rows, frames=len(c), branches=branches, ref=ref,
                                      shard=self.shard_group if sharded else None))
        self._fill_conditioning(pose_fea, audio, full_mask, face_mask, lip_mask, encoder_hidden_states, first=True)
        self._signature = self._conditioning_signature(latents, pose_fea, audio, full_mask, encoder_hidden_states)
        self._owns_shard_group = need_group and self.shard_group is not None and self._made_group
        return self

    def close(self):
        """Release the peer buffers of a frame-shard group this loop created in prepare() (collective over the ranks)."""
        self._graph = None
        if getattr(self, "_owns_shard_group", False) and self.shard_group is not None:
            self.shard_group.close(self.group)
            self.shard_group, self._owns_shard_group = None, False

    def _project_banks(self):
        """Project the reference banks of all spatial blocks (eagerly, outside any graph) and return the storage
        pointers a captured graph depends on."""
        return tuple(b.bank_kv_storage() if b.bank_kv(self.eng) is not None else None for b in self.unet.spatial_blocks())

    @staticmethod
    def _conditioning_signature(latents, pose_fea, audio, full_mask, ehs):
        return (tuple(latents.shape), None if pose_fea is None else tuple(pose_fea.shape), tuple(audio.shape),
                tuple(tuple(m.shape) for m in full_mask), tuple(ehs.shape))

    def _fill_conditioning(self, pose_fea, audio, full_mask, face_mask, lip_mask, encoder_hidden_states, first: bool):
        """Whole-video constants -> per-unit tensors in kernel layout.  ``first`` allocates them; later calls write into
        the SAME storage, which is what a captured CUDA graph reads."""
        eng, dev, L, nb = self.eng, self.eng.device, self.L, self.nb
        pose_tok = eng.ncfhw_to_tokens(pose_fea) if pose_fea is not None else None      # (L,h,w,320)
        audio_all = audio.to(device=dev, dtype=eng.dtype).contiguous()                   # (nb,L,M,768)
        ehs = encoder_hidden_states.to(dev)
        masks = [[m.to(device=dev, dtype=torch.float32).contiguous() for m in ms] for ms in (full_mask, face_mask, lip_mask)]
        mask_pads = [[_pad16(m) for m in ms] for ms in masks]
        for e in self.prepared:
            nbr, fr, rows = len(e["branches"]), e["frames"], e["rows"]
            aud = eng.gather_rows(audio_all.view(nb * L, -1), rows, out=None if first else e["audio"].view(nbr * fr, -1))
            mk = [[eng.gather_rows(mp, rows)[:, : m.shape[1]].contiguous() for mp, m in zip(mps, ms)]
                  for mps, ms in zip(mask_pads, masks)]
            pose = None
            if pose_tok is not None:
                pose = eng.gather_rows(pose_tok, e["x_idx"], out=None if first else e["pose"])
            sel = ehs[list(e["branches"])].contiguous()
            if first:
                e.update(pose=pose, audio=aud.view(nbr, fr, audio_all.shape[2], audio_all.shape[3]), masks=mk, ehs=sel)
            else:
                for dst_r, src_r in zip(e["masks"], mk):
                    for dst, src in zip(dst_r, src_r):
                        dst.copy_(src)
                e["ehs"].copy_(sel)

    def reload(self, latents, pose_fea, audio, full_mask, face_mask, lip_mask, encoder_hidden_states):
        """The next video of the same shape: refill the latents and the per-unit conditioning in place, so the CUDA
        graph captured for the first video (and the peer buffers of a frame-shard group) keep serving.  The reference
        banks currently attached to the UNet (ReferenceAttentionControl.update / set_banks for the NEW reference image)
        are re-projected into the K/V buffers the graph reads; a bank whose shape changed cannot be served by the
        captured graph and raises."""
        sig = self._conditioning_signature(latents, pose_fea, audio, full_mask, encoder_hidden_states)
        if sig != self._signature:
            raise ValueError(f"reload() needs the shapes prepare() saw: {self._signature}, got {sig}")
        self.latents.copy_(latents.to(torch.float32))
        self._fill_conditioning(pose_fea, audio, full_mask, face_mask, lip_mask, encoder_hidden_states, first=False)
        ptrs = self._project_banks()
        if self._graph is not None and ptrs != self._bank_ptrs:
            raise ValueError("reload(): the reference banks changed shape (or appeared / disappeared) since the CUDA graph "
                             "was captured; build a new DenoiseLoop for this video")
        return self

    # ------------------------------------------------------------------ the hot loop
    def _forward_units(self, lat_tok):
        """This rank's (window, CFG-branch) forwards of one step.  Frame-sharded units run one by one (their motion modules
        exchange rows with the peers); the others go through ``forward_tokens_group``: the 64x64 level unit by unit, the
        32x32 / 16x16 / 8x8 levels of ALL units as one batch (``deep_batch`` units, from down block ``deep_from`` on, or
        ``level_batch`` units per level)."""
        eng, u = self.eng, self.unet

        def finish(e, out):
            nbr, F_ = len(e["branches"]), e["frames"]
            pred = eng.tokens_to_ncfhw(out, nbr, F_, torch.float32)
            eng.window_accumulate(self.noise_acc, pred, e["idx"], e["branches"][0])
        plain = []
        for e in self.prepared:
            x = eng.gather_rows(lat_tok, e["x_idx"])
            nbr, F_ = len(e["branches"]), e["frames"]
            if e["shard"] is not None:
                finish(e, u.forward_tokens(eng, x, self.t_dev, e["ehs"], e["audio"], e["pose"], e["masks"][0], e["masks"][1],
                                           e["masks"][2], self.motion_scale, nbr, F_, ref_index=e["ref"], shard=e["shard"],
                                           time_proj=self._time_proj))
            else:
                plain.append((e, dict(x=x, pose=e["pose"], ehs=e["ehs"], audio=e["audio"], full=e["masks"][0],
                                      face=e["masks"][1], body=e["masks"][2], B=nbr, F=F_, ref=e["ref"])))
        if not plain:
            return
        same_f = len({un["F"] for _, un in plain}) == 1
        chunk = max(1, int(self.deep_batch)) * (2 if self.split_branches else 1) if same_f else 1
        for i in range(0, len(plain), chunk):
            part = plain[i:i + chunk]
            outs = u.forward_tokens_group(eng, [un for _, un in part], self.t_dev, self.motion_scale, time_proj=self._time_proj,
                                          deep_from=int(self.deep_from), level_batch=self.level_batch)
            for (e, _), out in zip(part, outs):
                finish(e, out)

    def _units_body(self):
        """Everything of a step that does not depend on host scalars: zero the accumulator, re-layout the latents,
        run this rank's (window, branch) forwards and scatter-add their predictions."""
        self.noise_acc.zero_()
        lat_tok = self.eng.ncfhw_to_tokens(self.latents)[: self.L]      # (L,h,w,4) run dtype
        # the timestep is the same for every window of a step: time embedding + all 22 time_emb_proj once per step
        self._time_proj = self.unet.time_projections(self.eng, self.t_dev, self.nb)
        self._forward_units(lat_tok)

    def capture_graph(self):
        """Capture ``_units_body`` (tens of thousands of launches) into one CUDA graph.  The timestep lives in a
        device tensor and the DDIM coefficients stay outside the graph, so one graph serves all steps."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):       # eager warm-up: builds weight packs, bank K/V, scratch buffers
            n0 = self.eng.ctx.launches()
            self._units_body()
            self.graph_launches = self.eng.ctx.launches() - n0
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._bank_ptrs = self._project_banks()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._units_body()
        return self

    def step(self, i: int):
        """One DDIM step over the whole video (all context windows, both CFG branches)."""
        eng = self.eng
        t = self.timesteps[i]
        self.t_dev.fill_(float(t))
        if self._graph is not None:
            self._graph.replay()
        else:
            self._units_body()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.noise_acc, group=self.group)
        cx, cv = self.schedule.step_coefficients(t, self.n_steps)
        eng.cfg_ddim_step(self.latents, self.noise_acc, self.inv_count, self.cfg, self.guidance, cx, cv)
        return self.latents

    def run(self, callback: Optional[Callable] = None, callback_steps: int = 1):
        for i in range(len(self.timesteps)):
            self.step(i)
            if callback is not None and i % callback_steps == 0:
                callback(i, self.timesteps[i], self.latents)
        if self.shard_group is not None:
            self.shard_group.check()      # a peer-barrier timeout means rows were missing: never return such latents
        return self.latents


def plan_rank(n_windows: int, n_branches: int, rank: int, world: int, k: int, shard_remainder: bool):
    """This rank's work for one step: [(window index, branches, frame-sharded?)], and whether the run needs peer buffers at
    all (the same answer on every rank: creating them is a collective).  Frame-sharded forwards come first: all ranks leave
    the per-step all-reduce together, so the k peers of a group meet at once."""
    group_idx = rank // k
    if shard_remainder and k > 1:
        whole, shared = plan_units_mixed(n_windows, n_branches, world, k)
        units = [(wi, b, True) for wi, b in shared[group_idx]] + [(wi, b, False) for wi, b in whole[rank]]
        return units, any(len(s) for s in shared)
    return [(wi, b, k > 1) for wi, b in plan_units(n_windows, n_branches, world // k)[group_idx]], k > 1


def plan_units(n_windows: int, n_branches: int, n_groups: int):
    """Deal the (window, CFG-branch) forwards of one step to ``n_groups`` rank groups, as large as they come: first whole
    windows (both CFG branches batched as one B=2 forward: twice the GEMM M), n_windows // n_groups to every group; the
    windows left over are split into single-branch forwards and dealt round; what still does not divide goes to the first
    groups.  Returns one list of (window index, branches) per group."""
    whole, left = _deal_large(n_windows, n_branches, n_groups)
    for i, u in enumerate(left):
        whole[i % n_groups].append(u)
    return whole


def plan_units_mixed(n_windows: int, n_branches: int, world: int, k: int):
    """Like plan_units over all ``world`` ranks, but the single-branch forwards that do not divide over the ranks are shared
    out to the world // k groups of k ranks, which run them frame-sharded (20 forwards on 8 GPUs: one B=2 window per rank,
    and each pair of ranks shares one more single-branch forward; when the leftovers do not divide over the groups either
    they are dealt whole to the first ranks).  Returns (whole: one list per rank, shared: one list per group)."""
    whole, left = _deal_large(n_windows, n_branches, world)
    groups = world // k
    shared = [[] for _ in range(groups)]
    if left and len(left) % groups == 0:
        for i, u in enumerate(left):
            shared[i % groups].append(u)
    else:
        for i, u in enumerate(left):
            whole[i % world].append(u)
    return whole, shared


def _deal_large(n_windows: int, n_branches: int, n: int):
    """-> (per-receiver lists after the even deals, the forwards left over: fewer than n single-branch ones)."""
    out = [[] for _ in range(n)]
    first_single = 0
    if n_branches == 2:
        per = n_windows // n
        for r in range(n):
            out[r] += [(wi, (0, 1)) for wi in range(r * per, (r + 1) * per)]
        first_single = per * n
    singles = [(wi, (b,)) for wi in range(first_single, n_windows) for b in range(n_branches)]
    per = len(singles) // n
    for r in range(n):
        out[r] += singles[r * per:(r + 1) * per]
    return out, singles[per * n:]


def _pad16(m: torch.Tensor) -> torch.Tensor:
    """Rows must be multiples of 16 bytes for mmgt_gather_rows (the 8x8 level has 64 floats: fine; 4x4 not)."""
    cols = m.shape[1]
    pad = (-cols) % 4
    if pad == 0:
        return m
    out = torch.zeros((m.shape[0], cols + pad), device=m.device, dtype=m.dtype)
    out[:, :cols] = m
    return out


class Pose2VideoPipeline:
    """Same constructor / ``__call__`` signature as the reference (pipeline_pose2vid_long.py:38-56,337-366), and the same
    argument types: ``ref_image`` a PIL image, ``pose_images`` a list of PIL images, masks as lists of 4 ``(L, T_l)``
    tensors, ``audio_tensor`` (1, L, 32, 768) -- so ``scripts/pose2vid.py:284-296`` / ``scripts/audio2vid.py:484-498``
    call it unchanged.  The one-shot conditioning passes run on the modules handed to the constructor (CLIP image
    encoder, VAE encoder, ReferenceNet, PoseGuider: the reference's PyTorch); the 30-step loop runs
    on the sm_100a kernels through a ``DenoiseLoop`` whose CUDA graph is cached on the pipeline per video shape, so a
    second call of the same shape only refills the conditioning (``DenoiseLoop.reload``)."""

    def __init__(self, vae, image_encoder, reference_unet, denoising_unet, pose_guider, scheduler, image_proj_model=None,
                 tokenizer=None, text_encoder=None):
        self.vae, self.image_encoder, self.reference_unet = vae, image_encoder, reference_unet
        self.denoising_unet, self.pose_guider, self.scheduler = denoising_unet, pose_guider, scheduler
        self.image_proj_model, self.tokenizer, self.text_encoder = image_proj_model, tokenizer, text_encoder
        boc = getattr(getattr(vae, "config", None), "block_out_channels", None)
        self.vae_scale_factor = 2 ** (len(boc) - 1) if boc is not None else 8
        self.ref_image_processor = VaeImageProcessor(self.vae_scale_factor, do_convert_rgb=True)
        self.cond_image_processor = VaeImageProcessor(self.vae_scale_factor, do_convert_rgb=True, do_normalize=False)
        self.rank, self.world_size, self.process_group = 0, 1, None     # set by the launcher for multi-GPU runs
        self.frame_shards = 1                                           # k ranks split the frames of a window
        self.shard_remainder = False                                    # only the forwards left over by the whole deal
        self.use_cuda_graph = True
        self.deep_batch = None                                          # None = DenoiseLoop's default (all units of a rank)
        self.deep_from = None                                           # None = DenoiseLoop's default (1)
        self.level_batch = None                                         # units per batch at every level (overrides deep_from)
        self.split_branches = None                                      # None = DenoiseLoop's default
        self.decode_chunk_size = 8                                      # frames per vae.decode call (the reference: 1)
        self.shard_decode = True                                        # multi-GPU: every rank decodes a slice of the frames
        self._loops = {}                                                # video shape -> captured DenoiseLoop

    def to(self, *a, **k):
        for m in (self.vae, self.image_encoder, self.reference_unet, self.denoising_unet, self.pose_guider):
            if m is not None and hasattr(m, "to"):
                m.to(*a, **k)
        return self

    @property
    def device(self):
        return self.denoising_unet.device

    @property
    def _execution_device(self):
        return self.denoising_unet.device

    def enable_vae_slicing(self):
        self.vae.enable_slicing()

    def disable_vae_slicing(self):
        self.vae.disable_slicing()

    def close(self):
        """Drop the cached loops (CUDA graphs, peer buffers)."""
        for loop in self._loops.values():
            loop.close()
        self._loops = {}

    def prepare_latents(self, batch_size, num_channels_latents, width, height, video_length, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_channels_latents, video_length, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective "
                             f"batch size of {batch_size}.")
        if latents is None:
            if isinstance(generator, list):
                generator = generator[0]
            gdev = generator.device if generator is not None else torch.device("cpu")
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to(device)
        else:
            latents = latents.to(device)
        return latents * 1.0   # init_noise_sigma == 1 for DDIM

    def decode_latents(self, latents, decode_chunk_size: Optional[int] = None):
        """VAE decode of the L frames (pipeline_pose2vid_long.py:112-125: the reference decodes ONE frame per ``vae.decode``
        call, 80 sequential 512x512 decodes).  The VAE is per-frame, so frames are decoded ``decode_chunk_size`` at a
        time (default ``self.decode_chunk_size``) and, with more than one rank (``self.shard_decode``), every rank decodes a
        contiguous slice of the frames and the slices are all-gathered -- same values, 1 / (chunk x ranks) of the
        sequential latency.  The VAE itself stays the caller's PyTorch module (north_star: out of the hot path).
        -> numpy (1, 3, L, H, W) float32 in [0, 1]."""
        chunk = int(decode_chunk_size or self.decode_chunk_size)
        L = latents.shape[2]
        lat = (1 / 0.18215 * latents)[0].permute(1, 0, 2, 3)                               # (L, 4, h, w) = "(b f) c h w"
        lo, hi = 0, L
        sharded = self.shard_decode and self.world_size > 1
        if sharded:
            per = (L + self.world_size - 1) // self.world_size
            lo, hi = min(L, self.rank * per), min(L, (self.rank + 1) * per)
        frames = [self.vae.decode(lat[i:min(hi, i + chunk)].to(self.vae.dtype)).sample for i in range(lo, hi, chunk)]
        video = torch.cat(frames) if frames else lat.new_zeros((0, 3, lat.shape[2] * self.vae_scale_factor,
                                                                  lat.shape[3] * self.vae_scale_factor))
        if sharded:
            import torch.distributed as dist
            per = (L + self.world_size - 1) // self.world_size
            pad = torch.zeros((per,) + tuple(video.shape[1:]), device=video.device, dtype=video.dtype)
            pad[: video.shape[0]] = video
            parts = [torch.empty_like(pad) for _ in range(self.world_size)]
            dist.all_gather(parts, pad, group=self.process_group)
            video = torch.cat(parts)[:L]
        video = video.permute(1, 0, 2, 3).unsqueeze(0)                                       # (1, 3, L, H, W)
        return ((video / 2 + 0.5).clamp(0, 1)).cpu().float().numpy()

    def interpolate_latents(self, latents: torch.Tensor, interpolation_factor: int, device=None):
        """pipeline_pose2vid_long.py:292-332 with the linear method of src/pipelines/utils.py:15-16."""
        if interpolation_factor < 2:
            return latents
        L = latents.shape[2]
        out = torch.zeros(latents.shape[:2] + ((L - 1) * interpolation_factor + 1,) + latents.shape[3:],
                          device=latents.device, dtype=latents.dtype)
        rate = [i / interpolation_factor for i in range(interpolation_factor)][1:]
        k = 0
        for i0 in range(L - 1):
            v0, v1 = latents[:, :, i0], latents[:, :, i0 + 1]
            out[:, :, k] = v0
            k += 1
            for f in rate:
                out[:, :, k] = (1.0 - f) * v0 + f * v1
                k += 1
        out[:, :, k] = latents[:, :, L - 1]
        return out

    # ------------------------------------------------------------------ one-shot conditioning (reference PyTorch modules)
    def _clip_embeds(self, ref_image, device):
        from transformers import CLIPImageProcessor
        if not hasattr(self, "clip_image_processor"):
            self.clip_image_processor = CLIPImageProcessor()
        clip_image = self.clip_image_processor.preprocess(ref_image.resize((224, 224)), return_tensors="pt").pixel_values
        return self.image_encoder(clip_image.to(device, dtype=self.image_encoder.dtype)).image_embeds

    def _reference_banks(self, reader, ref_image, ehs, width, height, cfg, device):
        """VAE-encode the reference image, run the ReferenceNet once in write mode, hand its banks to the reader
        (pipeline_pose2vid_long.py:403-448,510-520)."""
        writer = ReferenceAttentionControl(self.reference_unet, do_classifier_free_guidance=cfg, mode="write", batch_size=1,
                                           fusion_blocks="full")
        try:
            ref = self.ref_image_processor.preprocess(ref_image, height=height, width=width)
            vdev = getattr(self.vae, "device", device)
            ref = ref.to(dtype=self.vae.dtype, device=vdev)
            ref_latents = self.vae.encode(ref).latent_dist.mean * 0.18215                  # (1, 4, h, w)
            rdt = getattr(self.reference_unet, "dtype", ref_latents.dtype)
            self.reference_unet(ref_latents.to(device=device, dtype=rdt).repeat(2 if cfg else 1, 1, 1, 1),
                                torch.zeros((), device=device, dtype=torch.long),
                                encoder_hidden_states=ehs.to(device=device, dtype=rdt), return_dict=False)
            reader.update(writer)
        finally:
            writer.clear()
            writer.remove()

    def _pose_features(self, pose_images, width, height, device):
        if torch.is_tensor(pose_images):               # already (1, 3, L, H, W) in [0, 1]
            cond = pose_images
        else:
            frames = [self.cond_image_processor.preprocess(p, height=height, width=width).unsqueeze(2) for p in pose_images]
            cond = torch.cat(frames, dim=2)                                                 # (1, 3, L, H, W)
        return self.pose_guider(cond.to(device=device, dtype=self.pose_guider.dtype))

    @torch.no_grad()
    def __call__(self, ref_image, pose_images, audio_tensor, pixel_values_full_mask, pixel_values_face_mask,
                 pixel_values_lip_mask, width, height, video_length, num_inference_steps, guidance_scale,
                 num_images_per_prompt=1, eta: float = 0.0, motion_scale=None, generator=None, output_type="tensor",
                 return_dict: bool = True, callback=None, callback_steps=1, context_schedule="uniform", context_frames=12,
                 context_stride=1, context_overlap=4, context_batch_size=1, interpolation_factor=1, **kwargs):
        """Extra keyword arguments (all optional, none used by the reference scripts): ``latents`` initial noise;
        ``clip_image_embeds`` / ``reference_banks`` / ``pose_fea`` precomputed conditioning that replaces the
        corresponding one-shot pass."""
        if eta != 0.0:
            raise NotImplementedError("eta != 0 (the reference scripts use the DDIM default eta = 0)")
        if num_images_per_prompt != 1:
            raise NotImplementedError("num_images_per_prompt != 1 (the reference hard-codes batch_size = 1)")
        sample_size = getattr(getattr(self.denoising_unet, "config", None), "sample_size", None)
        height = height or sample_size * self.vae_scale_factor
        width = width or sample_size * self.vae_scale_factor
        device = self._execution_device
        cfg = guidance_scale > 1.0
        # context_batch_size only batches independent windows in the reference (:546-552); windows run one by one here
        # with identical results, so the value is accepted and ignored.
        # --- CLIP image embedding (:380-394)
        clip_embeds = kwargs.get("clip_image_embeds")
        if clip_embeds is None:
            clip_embeds = self._clip_embeds(ref_image, device)
        ehs = clip_embeds.unsqueeze(1) if clip_embeds.dim() == 2 else clip_embeds
        if cfg:
            ehs = torch.cat([torch.zeros_like(ehs), ehs], dim=0)
        # --- reference features (:396-409, 439-448, 510-520)
        reader = ReferenceAttentionControl(self.denoising_unet, do_classifier_free_guidance=cfg, mode="read", batch_size=1,
                                           fusion_blocks="full")
        banks = kwargs.get("reference_banks")
        if banks is not None:
            reader.set_banks(banks)
        elif kwargs.get("reference_control_writer") is not None:      # a writer the caller already ran
            reader.update(kwargs["reference_control_writer"])
        else:
            self._reference_banks(reader, ref_image, ehs, width, height, cfg, device)
        latents = self.prepare_latents(num_images_per_prompt, self.denoising_unet.in_channels, width, height, video_length,
                                       torch.float32, device, generator, kwargs.get("latents"))
        # --- pose features (:450-464)
        pose_fea = kwargs.get("pose_fea")
        if pose_fea is None:
            pose_fea = self._pose_features(pose_images, width, height, device)
        # --- masks and audio, duplicated for CFG (:466-488)
        dup = (lambda ms: [torch.cat([m] * 2) for m in ms]) if cfg else (lambda ms: list(ms))
        audio = audio_tensor.to(device)
        if cfg:
            audio = torch.cat([torch.zeros_like(audio), audio], dim=0)
        full, face, lip = dup(pixel_values_full_mask), dup(pixel_values_face_mask), dup(pixel_values_lip_mask)
        # --- the denoising loop (:491-646): cached per video shape
        sched = self.scheduler if isinstance(self.scheduler, DDIMSchedule) else DDIMSchedule.from_scheduler(self.scheduler)
        ms_key = None if motion_scale is None else tuple(float(m) for m in motion_scale)
        key = (tuple(latents.shape), num_inference_steps, float(guidance_scale), context_schedule, context_frames,
               context_stride, context_overlap, ms_key, self.rank, self.world_size, self.frame_shards, self.shard_remainder,
               tuple(audio.shape), str(self.denoising_unet._engine(device).dtype))
        loop = self._loops.get(key)
        try:
            if loop is None:
                loop = DenoiseLoop(self.denoising_unet, sched, num_inference_steps, guidance_scale, context_frames,
                                   context_stride, context_overlap, context_schedule, motion_scale, self.rank, self.world_size,
                                   self.process_group, self.frame_shards, shard_remainder=self.shard_remainder)
                if self.deep_batch is not None:
                    loop.deep_batch = int(self.deep_batch)
                if self.deep_from is not None:
                    loop.deep_from = int(self.deep_from)
                if self.level_batch is not None:
                    loop.level_batch = [int(v) for v in self.level_batch]
                if self.split_branches is not None:
                    loop.split_branches = bool(self.split_branches)
                loop.prepare(latents, pose_fea, audio, full, face, lip, ehs)
                if self.use_cuda_graph and latents.is_cuda:
                    loop.capture_graph()
                self._loops[key] = loop
            else:
                loop.reload(latents, pose_fea, audio, full, face, lip, ehs)
            latents = loop.run(callback, callback_steps).clone()
        except Exception:
            bad = self._loops.pop(key, None)
            if bad is not None:
                bad.close()
            raise
        finally:
            reader.clear()
        if interpolation_factor > 0:
            latents = self.interpolate_latents(latents, interpolation_factor, device)
        if output_type == "latent" or self.vae is None:
            images = latents
        else:
            images = self.decode_latents(latents)
            if output_type == "tensor":
                images = torch.from_numpy(images)
        if not return_dict:
            return images
        return Pose2VideoPipelineOutput(videos=images)
