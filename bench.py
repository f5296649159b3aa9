#!/usr/bin/env python
"""bench.py -- MMGT stage-2 denoise throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our sm_100a path (torchrun launches N ranks for N > 1)
  python bench.py --impl reference --steps K --warmup W  the reference algorithm on the host CPU cores (oracle port)

Workload (BASELINE.json configs[1]): pose2vid 512x512 -> 64x64 latent, 80 frames = 10 context windows of 12
frames, CFG 3.5, 30 DDIM steps, random-init full-width UNet3D (1.40 G parameters), synthetic inputs.
A "step" is ONE DDIM step over the whole video: 10 windows x 2 CFG branches = 20 UNet window forwards,
overlap-average, CFG combine, DDIM update.  frames/s = 80 / (30 * seconds_per_step).
Multi-GPU: the 20 (window, branch) forwards of a step are dealt to the ranks, one all-reduce (NCCL) of the
accumulated prediction per step => strong scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

VIDEO_LENGTH, LATENT, N_STEPS, GUIDANCE = 80, 64, 30, 3.5
FLOP_PER_UNIT = 31.595e12 / 2       # algorithmic FLOPs of one (window, CFG-branch) forward, SURVEY section 8d (B=2 window / 2)
UNITS_PER_STEP = 20
# BASELINE.json configs (1-based like SURVEY section 8d).  2 / 3 = the headline (N = 1 / N > 1); 4 = one independent
# audio2vid clip per GPU (replicas, non-zero audio); 5 = 768x768, 160 frames (20 windows, 86.67 TFLOP per window forward).
CONFIGS = {2: dict(frames=80, latent=64, flop_per_unit=31.595e12 / 2, name="pose2vid 512x512 (64x64 latent), 80 frames"),
           4: dict(frames=80, latent=64, flop_per_unit=31.595e12 / 2,
                   name="audio2vid 512x512 (64x64 latent), one independent 80-frame clip per GPU, non-zero audio tokens"),
           5: dict(frames=160, latent=96, flop_per_unit=86.67e12 / 2, name="long video 768x768 (96x96 latent), 160 frames")}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(n)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------ synthetic workload
def build_unet(device, compute_dtype):
    from mmgt_b200.unet_3d import UNet3DConditionModel
    cfg = dict(sample_size=64, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True, freq_shift=0,
               block_out_channels=[320, 640, 1280, 1280], layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1,
               act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768, attention_head_dim=8)
    extra = dict(use_inflated_groupnorm=True, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False,
                 use_motion_module=True, use_audio_module=True, motion_module_resolutions=[1, 2, 4, 8],
                 motion_module_mid_block=True, motion_module_decoder_only=False, motion_module_type="Vanilla",
                 motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                                           attention_block_types=["Temporal_Self", "Temporal_Self"],
                                           temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                           temporal_attention_dim_div=1),
                 audio_attention_dim=768, stack_enable_blocks_name=["up", "down", "mid"], stack_enable_blocks_depth=[0, 1, 2, 3])
    torch.manual_seed(0)
    unet = UNet3DConditionModel.from_config(cfg, **extra)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():   # the reference zero-initialises these; make MM-HAA / motion modules numerically live
        for n, p in unet.named_parameters():
            if "zero_conv" in n or "temporal_transformer.proj_out" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    unet.to(device)
    unet.set_compute_dtype(compute_dtype)
    unet.train()
    unet.enable_gradient_checkpointing()   # what scripts/pose2vid.py does => motion_scale reaches MM-HAA
    return unet


def synthetic_video(L, latent, pinned=True, seed=42):
    """Host-side (pinned) whole-video inputs in the layout Pose2VideoPipeline holds before the loop."""
    g = torch.Generator().manual_seed(seed)
    pin = (lambda t: t.pin_memory()) if pinned and torch.cuda.is_available() else (lambda t: t)
    latents = pin(torch.randn(1, 4, L, latent, latent, generator=g))
    clip = torch.randn(1, 1, 768, generator=g)
    ehs = pin(torch.cat([torch.zeros_like(clip), clip]))
    aud = torch.nn.functional.layer_norm(torch.randn(1, L, 32, 768, generator=g), (768,))
    audio = pin(torch.cat([torch.zeros_like(aud), aud]))
    pose = pin(0.1 * torch.randn(1, 320, L, latent, latent, generator=g))
    masks = []
    for k in range(3):
        lv = []
        for lvl in range(4):
            t = (latent >> lvl) ** 2
            m = torch.rand(L, t, generator=g)
            lv.append(pin(torch.cat([m, m]) + (1.0 if k == 0 else 0.0)))
        masks.append(lv)
    banks = []
    for c, lvl, cnt in ((1280, 2, 2), (1280, 2, 3), (1280, 3, 1), (640, 1, 2), (640, 1, 3), (320, 0, 2), (320, 0, 3)):
        for _ in range(cnt):
            banks.append(torch.randn(2, (latent >> lvl) ** 2, c, generator=g).half())
    return dict(latents=latents, ehs=ehs, audio=audio, pose=pose, full=masks[0], face=masks[1], lip=masks[2], banks=banks)


def to_device(v, dev):
    out = {}
    for k, x in v.items():
        out[k] = [t.to(dev, non_blocking=True) for t in x] if isinstance(x, list) else x.to(dev, non_blocking=True)
    return out


def h2d_bytes(v):
    tot = 0
    for k, x in v.items():
        if k == "banks":
            continue
        tot += sum(t.numel() * t.element_size() for t in x) if isinstance(x, list) else x.numel() * x.element_size()
    return tot


def _describe(world, shards, remainder, loop):
    whole = sum(1 for _, _, sh in loop.units if not sh)
    shared = sum(1 for _, _, sh in loop.units if sh)
    if shards == 1 or shared == 0:
        return f"(window,cfg-branch) forwards dealt whole over {world} rank(s)"
    if remainder:
        return (f"(window,cfg-branch) forwards over {world} ranks: {whole} whole unit(s) per rank + {shared} shared by each "
                f"group of {shards} ranks as frame shards (motion modules: peer-store row exchange over NVLink)")
    return (f"(window,cfg-branch) forwards over {world // shards} rank group(s) x {shards} frame shards per window "
            "(motion modules: peer-store row exchange over NVLink)")


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from mmgt_b200.pipeline_pose2vid_long import DenoiseLoop, Pose2VideoPipeline
    from mmgt_b200.scheduling_ddim import DDIMSchedule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    cdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    conf = CONFIGS[args.config]
    L = args.frames or conf["frames"]
    latent = conf["latent"]
    replicas = args.config == 4            # one independent clip per GPU: no data-path collective at all
    t0 = time.time()
    unet = build_unet(dev, cdt)
    host = synthetic_video(L, latent, seed=42 + (rank if replicas else 0))
    host2 = synthetic_video(L, latent, seed=1042 + (rank if replicas else 0))    # the NEXT video (another reference image)
    log(f"[rank {rank}] model + inputs built in {time.time() - t0:.1f}s")
    sched = DDIMSchedule.from_config()

    # Schedule.  Default ("auto"): whole windows (B=2), then single-branch forwards, dealt evenly to the ranks; forwards that
    # still do not divide (10 windows on 8 GPUs: one window each, 4 forwards left) are shared by pairs of ranks as frame
    # shards (2.5 forwards of work per rank instead of 3 / 2).
    loop_world, loop_rank = (1, 0) if replicas else (world, rank)
    shards, remainder = args.frame_shards, args.shard_remainder
    from mmgt_b200.context import get_context_scheduler
    n_windows = len(list(get_context_scheduler("uniform")(0, N_STEPS, L, 12, 1, 4)))
    units_per_step = 2 * n_windows
    if shards == 0:
        from mmgt_b200.pipeline_pose2vid_long import plan_units_mixed
        shards, remainder = 1, False
        if loop_world % 2 == 0 and any(plan_units_mixed(n_windows, 2, loop_world, 2)[1]):
            shards, remainder = 2, True
    if loop_world % shards:
        raise SystemExit(f"bench.py: --frame-shards {shards} must divide the number of ranks {loop_world}")

    # Everything goes through the reference-facing plugin call, Pose2VideoPipeline.__call__ (output_type="latent": the VAE
    # is out of the hot path).  The one-shot conditioning passes (CLIP, VAE encode, ReferenceNet, PoseGuider) are outside the
    # metric: their results are handed in as HOST tensors through the optional keyword arguments.
    pipe = Pose2VideoPipeline(vae=None, image_encoder=None, reference_unet=None, denoising_unet=unet, pose_guider=None,
                              scheduler=sched)
    pipe.rank, pipe.world_size, pipe.process_group = loop_rank, loop_world, None
    pipe.frame_shards, pipe.shard_remainder = shards, remainder
    pipe.use_cuda_graph = not args.no_graph
    pipe.deep_batch = args.deep_batch
    pipe.deep_from = args.deep_from
    pipe.level_batch = [int(v) for v in args.level_batch.split(",")] if args.level_batch else None
    pipe.split_branches = None if args.split_branches is None else bool(args.split_branches)
    eng = unet._engine(dev)
    if args.no_tc:
        eng.ctx.set_tensor_cores(False)
    if args.strict:
        eng.ctx.set_strict_tensor_cores(True)
    eng.unfused_exchange = args.unfused_exchange
    if args.pdl is not None:
        eng.ctx.set_pdl(bool(args.pdl))
    for name, setter in (("gn_split", eng.ctx.set_groupnorm_split), ("conv_implicit", eng.ctx.set_conv_implicit_all),
                         ("geglu_exact", eng.ctx.set_geglu_exact), ("attn_v2", eng.ctx.set_attention_v2), ("attn_persist", eng.ctx.set_attention_persistent),
                         ("attn_q256", eng.ctx.set_attention_q256), ("attn_packed", eng.ctx.set_attention_packed), ("ln_persist", eng.ctx.set_layernorm_persistent),
                         ("temporal_rows", eng.ctx.set_temporal_rows), ("lean_epilogue", eng.ctx.set_lean_epilogue),
                         ("tma_store", eng.ctx.set_tma_store), ("residual_mma", eng.ctx.set_residual_mma)):
        v = getattr(args, name)
        if v is not None:
            setter(int(v))
    if args.fuse_ln is not None:
        eng.fuse_layernorm = bool(args.fuse_ln)

    def call_pipeline(h, steps, callback=None):
        cond = lambda ms: [m[:L] for m in ms]   # noqa: E731  (the pipeline duplicates the masks for CFG itself)
        return pipe(ref_image=None, pose_images=None, audio_tensor=h["audio"][1:2], pixel_values_full_mask=cond(h["full"]),
                    pixel_values_face_mask=cond(h["face"]), pixel_values_lip_mask=cond(h["lip"]), width=latent * 8,
                    height=latent * 8, video_length=L, num_inference_steps=steps, guidance_scale=GUIDANCE,
                    motion_scale=[1.0, 1.0, 2.0], output_type="latent", callback=callback, callback_steps=1,
                    clip_image_embeds=h["ehs"][1], reference_banks=h["banks"], pose_fea=h["pose"], latents=h["latents"]).videos

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- first video: builds the weight packs, projects the reference banks, captures the CUDA graph of one DDIM step
    torch.cuda.synchronize()
    t_first0 = time.perf_counter()
    for attempt in range(2):
        try:
            first = call_pipeline(host, 2 if (args.ncu_step or args.quick) else N_STEPS)   # (profiling / A-B runs only need the captured loop)
            break
        except RuntimeError as e:
            if attempt == 0 and args.frame_shards == 0 and "peer memory" in str(e):   # raised on every rank together
                log(f"[rank {rank}] {e}; falling back to whole forwards only")
                pipe.frame_shards, pipe.shard_remainder = 1, False
                shards, remainder = 1, False
                continue
            raise
    torch.cuda.synchronize()
    first_video_s = time.perf_counter() - t_first0      # one-off per process and video shape (packs, graph capture) + 30 steps
    loop = next(iter(pipe._loops.values()))
    assert torch.isfinite(first).all(), "non-finite latents after the first video"
    log(f"[rank {rank}] first video through Pose2VideoPipeline.__call__: {first_video_s:.1f}s "
        f"({loop.graph_launches} kernel launches per DDIM step in the graph)")

    # ---- device-resident timing: W warm-up + K timed steps, CUDA events, max over ranks
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl
    reader = ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")

    def attach_banks(h):     # (the pipeline call clears the reference banks when it returns, like the reference)
        reader.set_banks([b.to(dev) for b in h["banks"]])
    d = to_device(host, dev)
    attach_banks(host)
    loop.reload(d["latents"], d["pose"], d["audio"], d["full"], d["face"], d["lip"], d["ehs"])
    for i in range(args.warmup):
        loop.step(i % len(loop.timesteps))
    barrier()
    if args.ncu_step:
        # `ncu --profile-from-start off --metrics gpu__time_duration.sum ... python bench.py --ncu-step`: exactly one DDIM
        # step (eager launches) between cudaProfilerStart / Stop -> the launch list under profiles/; prints no bench line.
        graph, loop._graph = loop._graph, None
        torch.cuda.cudart().cudaProfilerStart()
        loop.step(args.warmup % len(loop.timesteps))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        loop._graph = graph
        log("[ncu-step] one step profiled; no bench line is printed under a profiler")
        return
    sampler = ClockSampler(local)
    sampler.start()
    n0, s0 = eng.ctx.launches(), eng.ctx.simt_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loop.step((args.warmup + i) % len(loop.timesteps))
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.ctx.launches() - n0 + (loop.graph_launches * args.steps if loop._graph is not None else 0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    n_videos = world if replicas else 1
    value = n_videos * L / (N_STEPS * ms_per_step / 1e3)
    if args.quick:      # A/B aid: the device-resident step time only (no e2e / parity / roofline / CPU legs)
        if rank == 0:
            print(json.dumps(dict(quick=True, ms_per_step=ms_per_step, value=value, n_gpus=world, config=args.config,
                                  flags=dict(gn_split=args.gn_split, conv_implicit=args.conv_implicit, geglu_exact=args.geglu_exact,
                                             fuse_ln=args.fuse_ln, attn_v2=args.attn_v2, attn_persist=args.attn_persist, attn_q256=args.attn_q256, attn_packed=args.attn_packed, ln_persist=args.ln_persist, deep_batch=args.deep_batch, deep_from=args.deep_from, level_batch=args.level_batch,
                                             split_branches=args.split_branches,
                                             temporal_rows=args.temporal_rows, lean_epilogue=args.lean_epilogue, tma_store=args.tma_store, residual_mma=args.residual_mma), simt_launches=eng.ctx.simt_launches() - s0,
                                  gpu_launches=launches, clocks=clocks)), flush=True)
        pipe.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity of the multi-GPU schedule, asserted in the run itself: two DDIM steps from the same latents through the
    #      distributed loop (every rank ends with the same latents after the all-reduce) and, on rank 0, through a local
    #      single-GPU loop that runs all 20 forwards of a step as whole B=2 windows.
    parity = None
    if world > 1 and not replicas:
        loop.latents.copy_(d["latents"])
        for i in (3, 17):
            loop.step(i)
        lat_dist = loop.latents.clone()
        if rank == 0:
            solo = DenoiseLoop(unet, sched, N_STEPS, GUIDANCE, motion_scale=[1.0, 1.0, 2.0])
            solo.prepare(d["latents"], d["pose"], d["audio"], d["full"], d["face"], d["lip"], d["ehs"])
            for i in (3, 17):
                solo.step(i)
            torch.cuda.synchronize()
            err = float((lat_dist.double() - solo.latents.double()).norm() / solo.latents.double().norm())
            parity = dict(latents_rel_l2_vs_n1=err, steps=2, tolerance=1e-2,
                          what="2 DDIM steps: this run's schedule over all ranks vs all (window, CFG-branch) forwards on rank 0 alone")
            del solo
        barrier()
        flag = torch.tensor([0.0 if parity is None or parity["latents_rel_l2_vs_n1"] < 1e-2 else 1.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag) != 0.0:
            raise SystemExit(f"bench.py: multi-GPU latents differ from the single-GPU result: {parity}")
    elif loop._graph is not None:
        # N = 1: the graph replay against the same two steps launched eagerly (must be bit-identical)
        loop.latents.copy_(d["latents"])
        for i in (3, 17):
            loop.step(i)
        lat_graph = loop.latents.clone()
        graph, loop._graph = loop._graph, None
        loop.latents.copy_(d["latents"])
        for i in (3, 17):
            loop.step(i)
        loop._graph = graph
        err = float((lat_graph.double() - loop.latents.double()).norm() / loop.latents.double().norm())
        parity = dict(latents_rel_l2_graph_vs_eager=err, steps=2,
                      vs_reference="tests/test_config2_gpu.py: this workload's shapes against the reference's own modules "
                                   "(tests/golden/config2_step.npz)")

    # ---- end-to-end through the plugin call with HOST buffers: a NEW video (other latents, reference banks, CLIP vector,
    #      pose, audio, masks, all in pinned host memory) through Pose2VideoPipeline.__call__.  Inside the timed region:
    #      the H2D upload of the conditioning and its re-layout into the loop's static buffers, the re-projection of the
    #      reference banks, all 30 DDIM steps (CUDA-graph replays), a D2H read of the latents after EVERY step (callback)
    #      and of the final result.
    out_host = torch.empty_like(host2["latents"]).pin_memory()
    n_cb = [0]

    def on_step(i, t, lat):
        out_host.copy_(lat, non_blocking=True)
        n_cb[0] += 1
    barrier()
    t_e0 = time.perf_counter()
    res = call_pipeline(host2, N_STEPS, callback=on_step)
    out_host.copy_(res, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_e0
    assert n_cb[0] == N_STEPS and len(pipe._loops) == 1, "the second video must reuse the captured loop"
    e2e_t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_videos * L / float(e2e_t)
    lat_bytes = host2["latents"].numel() * 4
    cond_bytes = h2d_bytes(host2) + sum(b.numel() * b.element_size() for b in host2["banks"])

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launch stream over one more step
    roof = None
    attach_banks(host2)
    loop.reload(d["latents"], d["pose"], d["audio"], d["full"], d["face"], d["lip"], d["ehs"])
    # every rank runs this extra eager step (it contains the per-step all-reduce); only rank 0 instruments it
    eng.prof = {} if rank == 0 else None
    graph, loop._graph = loop._graph, None        # per-launch events need eager launches
    loop.step(0)
    loop._graph = graph
    barrier()
    if rank == 0:
        prof, eng.prof = eng.prof, None
        rows = []
        for key, r in prof.items():
            tms = sum(a.elapsed_time(b) for a, b in r["events"])
            rows.append((tms, key, r))
        rows.sort(reverse=True, key=lambda x: x[0])
        tot = sum(x[0] for x in rows)
        pk = peaks()
        log(f"--- per-operator time over one step (event-timed, {tot:.1f} ms in instrumented ops) ---")
        table = []
        for tms, key, r in rows:
            tf = r["flops"] / (tms * 1e-3) / 1e12 if tms > 0 else 0
            gb = r["bytes"] / (tms * 1e-3) / 1e9 if tms > 0 else 0
            table.append(f"{tms:9.2f} ms {100 * tms / tot:5.1f}%  calls {r['calls']:5d}  {tf:8.1f} TFLOP/s {gb:8.1f} GB/s  {key}")
        for line in table[:40]:
            log(line)
        if args.ops_out:
            with open(args.ops_out, "w") as f:
                f.write(f"# per-operator CUDA-event time over one eager DDIM step ({tot:.1f} ms in instrumented ops); "
                        f"step under CUDA graph: {ms_per_step:.1f} ms\n" + "\n".join(table) + "\n")
        tms, key, r = rows[0]
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)["per_launch_bytes"].get(str(key))
        except Exception:
            traffic = None
        if r["flops"] > 0 and key[0] in ("gemm", "conv3x3", "attention"):
            ach = r["flops"] / (tms * 1e-3) / 1e12
            roof = dict(bound="tensor", kernel=str(key), achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s",
                        frac=ach / pk["tf_sustained"], traffic=traffic, peak_source=pk["src"] + " sustained bf16",
                        launches=r["calls"], avg_launch_ms=tms / r["calls"], share_of_step=tms / tot)
        else:
            ach = r["bytes"] / (tms * 1e-3) / 1e9
            roof = dict(bound="hbm", kernel=str(key), achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                        traffic=traffic, peak_source=pk["src"], launches=r["calls"], avg_launch_ms=tms / r["calls"],
                        share_of_step=tms / tot)
        whole = units_per_step * conf["flop_per_unit"] * n_videos / (ms_per_step * 1e-3) / 1e12 / world
        roof["whole_step_tflops_per_gpu"] = whole
        roof["whole_step_frac_of_peak"] = whole / pk["tf_sustained"]

    if loop.shard_group is not None:
        loop.shard_group.check()                   # raises if any peer barrier timed out during the run
    simt = eng.ctx.simt_launches() - s0
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 2:
        cpu = cpu_reference(unet, steps=1, warmup=0, budget_s=args.cpu_budget)

    if rank == 0:
        metric = "UNet3D denoise frames/s @512x512 80f (30-step CFG DDIM, 10 windows x 12 frames)"
        if args.config != 2:
            metric = f"UNet3D denoise frames/s, BASELINE config {args.config}: {conf['name']} (30-step CFG DDIM, {n_windows} windows x 12 frames)"
        line = dict(metric=metric, value=value,
                    unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="weak" if replicas else "strong", vs_baseline=None, dtype=args.dtype,
                    data="synthetic",
                    config=dict(workload=f"{conf['name']}, {L} frames, 30 DDIM steps, CFG 3.5, full-width UNet3D random-init, "
                                         f"non-zero audio tokens and three motion masks; step = 1 DDIM step = {n_windows} "
                                         f"windows x 2 CFG branches" + (" per clip, one clip per GPU" if replicas else ""),
                                baseline_config=args.config,
                                parallelism=("independent replicas: one clip per GPU, no collective" if replicas else
                                             _describe(world, shards, remainder, loop)),
                                l2_policy="per-step working set (weights 2.8 GB + activations) exceeds the 126 MB L2",
                                tensor_cores=not args.no_tc, strict_tensor_cores=bool(args.strict),
                                programmatic_dependent_launch=eng.ctx.pdl(), layernorm_in_gemm_epilogue=eng.ln_fused,
                                psnr_note="parity / PSNR are asserted on LATENTS (tests/): the VAE is outside the hot path"),
                    clocks=clocks, gpu_launches=launches, simt_launches=simt,
                    e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=(lat_bytes + cond_bytes) // N_STEPS,
                             d2h_bytes_per_step=lat_bytes + lat_bytes // N_STEPS, video_s=float(e2e_t),
                             through="Pose2VideoPipeline.__call__(output_type='latent') on a new video with pinned host "
                                     "inputs: conditioning upload + reload + 30 graph-replayed DDIM steps + per-step D2H",
                             first_video_s=first_video_s),
                    parity=parity, roofline=roof, cpu_baseline=cpu,
                    frame_evals_per_s=n_videos * units_per_step * 12 / (ms_per_step / 1e3))
        print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_reference(unet_or_none, steps, warmup, budget_s):
    """Times the reference algorithm on the host cores (all of them, float32): the reference's OWN modules
    (/root/reference/src/models/*.py imported unchanged through oracle/reference_loader.py) where the checkout exists
    (kind "reference"), else the oracle port (kind "port": the GPU box has no /root/reference).
    One sample = one (window, CFG-branch) forward of config 2: B=1, F=12, 64x64 latent; 600 of them make a video."""
    from oracle import reference_loader as RL
    from oracle.unet3d import UNetSpec, bank_pairing_order, unet3d_forward, spatial_block_prefixes, spatial_block_width
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = UNetSpec()
    src = unet_or_none if unet_or_none is not None else build_unet(torch.device("cpu"), torch.float32)
    sd = {k: v.detach().float().cpu() for k, v in src.state_dict().items()}
    g = torch.Generator().manual_seed(42)
    Fr, lat = 12, LATENT
    sample = torch.randn(1, 4, Fr, lat, lat, generator=g)
    ehs = torch.randn(1, 1, 768, generator=g)
    aud = torch.nn.functional.layer_norm(torch.randn(1, Fr, 32, 768, generator=g), (768,))
    pose = 0.1 * torch.randn(1, 320, Fr, lat, lat, generator=g)
    masks = [[torch.rand(Fr, (lat >> l) ** 2, generator=g) for l in range(4)] for _ in range(3)]
    banks = {}
    for pre in spatial_block_prefixes(spec):
        c = spatial_block_width(spec, pre)
        lvl = {320: 0, 640: 1}.get(c, 3 if pre.startswith("mid") else 2)
        banks[pre] = torch.randn(2, (lat >> lvl) ** 2, c, generator=g).half().float()
    kind = "port"

    def unit():
        return unet3d_forward(sd, spec, sample, 500, ehs, aud, pose, masks[0], masks[1], masks[2], [1.0, 1.0, 2.0], banks,
                              ref_index=[1], apply_motion_scale=True)
    if RL.reference_available():
        try:
            ref_unet, mods = RL.build_reference_unet()
            ref_unet.load_state_dict(sd, strict=True)
            msa, attn_mod = mods["mutual_self_attention"], mods["attention"]
            msa.ReferenceAttentionControl(ref_unet, do_classifier_free_guidance=False, mode="read", batch_size=1,
                                          fusion_blocks="full")
            blocks = sorted([m for m in msa.torch_dfs(ref_unet) if isinstance(m, attn_mod.TemporalBasicTransformerBlock)],
                            key=lambda m: -m.norm1.normalized_shape[0])
            for blk, pre in zip(blocks, bank_pairing_order(spec)):
                blk.bank = [banks[pre][1:2].clone()]
            ref_unet.train()
            ref_unet.enable_gradient_checkpointing()        # the scripts' branch (scripts/pose2vid.py:151-156,183-184)
            kind = "reference"

            def unit():   # noqa: F811
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    return ref_unet(sample, torch.tensor(500), encoder_hidden_states=ehs, audio_embedding=aud, pose_cond_fea=pose,
                                    full_mask=masks[0], face_mask=masks[1], body_mask=masks[2], motion_scale=[1.0, 1.0, 2.0],
                                    return_dict=False)[0]
        except Exception as e:   # noqa: BLE001
            log(f"[cpu] reference modules not usable here ({type(e).__name__}: {e}); timing the oracle port")
            kind = "port"
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            unit()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            log(f"[cpu:{kind}] unit forward {i}: {dt:.1f}s")
            if time.perf_counter() - t_start > budget_s and times:
                break
    t_unit = sum(times) / len(times)
    units_per_video = N_STEPS * UNITS_PER_STEP
    return dict(value=VIDEO_LENGTH / (units_per_video * t_unit), unit="frames/s", cores=cores, kind=kind,
                sample=f"{len(times)} x one (window, CFG-branch) UNet3D forward of the same workload (B=1, 12 frames, 64x64 "
                       f"latent, fp32) = 1/{units_per_video} of the 30-step video, extrapolated; {t_unit:.1f}s each",
                seconds_per_unit=t_unit, steps_done=len(times))


def run_reference(args):
    """The CPU arm.  One STEP of this arm = one bounded sample = ONE (window, CFG-branch) UNet3D forward (1/20 of a DDIM
    step of the workload, 1/600 of the video): `steps` and `ms_per_step` are both in that unit, `value` extrapolates
    frames/s = 80 / (600 x seconds per sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_reference(None, steps=args.steps, warmup=min(args.warmup, 1), budget_s=args.cpu_budget)
    line = dict(impl="reference", metric="UNet3D denoise frames/s @512x512 80f (30-step CFG DDIM, 10 windows x 12 frames)",
                value=cpu["value"], unit="frames/s", n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=cpu["steps_done"],
                warmup=min(args.warmup, 1), ms_per_step=cpu["seconds_per_unit"] * 1e3, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="pose2vid 512x512 (64x64 latent), 80 frames, 30 DDIM steps, CFG 3.5, full-width UNet3D "
                                     "random-init; CPU arm: one step = one bounded sample = one (window, CFG-branch) forward "
                                     "= 1/20 of a DDIM step; value extrapolates x600",
                            step_is="one (window, CFG-branch) UNet3D forward (1/20 DDIM step)",
                            ddim_step_ms_extrapolated=cpu["seconds_per_unit"] * UNITS_PER_STEP * 1e3),
                cpu_baseline=cpu, e2e=dict(value=cpu["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--frames", type=int, default=0, help="video length (default: the config's)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="BASELINE.json config (1-based): 2 = headline (3 = the same at N > 1), 4 = one independent audio2vid "
                         "clip per GPU, 5 = 768x768 / 160 frames")
    ap.add_argument("--quick", action="store_true", help="A/B aid: print only the device-resident step time")
    ap.add_argument("--strict", action="store_true",
                    help="strict tensor-core mode: a bf16 operator without a tcgen05 kernel is an error, not a CUDA-core fallback")
    ap.add_argument("--gn-split", type=int, default=None, help="A/B: 1 / 0 two-kernel / fused spin-barrier GroupNorm")
    ap.add_argument("--conv-implicit", type=int, default=None, help="A/B: 1 / 0 implicit-GEMM / im2col stride-2 + upsample convs")
    ap.add_argument("--geglu-exact", type=int, default=None, help="A/B: 1 = erf GELU in the GEGLU epilogue")
    ap.add_argument("--attn-v2", type=int, default=None, help="A/B: 1 / 0 three-S-buffer / round-1 attention kernel (head dim <= 64)")
    ap.add_argument("--attn-q256", type=int, default=None, help="A/B: 1 / 0 256-query / 128-query CTAs in the head dim <= 64 attention kernel")
    ap.add_argument("--ln-persist", type=int, default=None, help="A/B: LayerNorm grid: 0 = 16 blocks per SM, 1 = one exact persistent wave, 2 = + prefetch")
    ap.add_argument("--attn-packed", type=int, default=None, help="A/B: 1 / 0 packed fp32 pairs (FFMA2 / FADD2) in the attention softmax loops")
    ap.add_argument("--attn-persist", type=int, default=None, help="A/B: 1 / 0 persistent / one-item-per-CTA attention kernel (head dim <= 64)")
    ap.add_argument("--deep-batch", type=int, default=None,
                    help="A/B: forwards whose 16x16 / 8x8 levels run as one batch (default: all of a rank's; 1 = unit by unit)")
    ap.add_argument("--deep-from", type=int, default=None,
                    help="A/B: first down block whose level runs batched over those forwards (default 2: 16x16 / 8x8; 1 adds 32x32)")
    ap.add_argument("--level-batch", default=None,
                    help="A/B: forwards per batch at each UNet level, e.g. 1,2,99,99 (overrides --deep-from)")
    ap.add_argument("--split-branches", type=int, default=None,
                    help="A/B: 1 = the CFG branches of a window are separate single-branch forwards (B = 1 units)")
    ap.add_argument("--residual-mma", type=int, default=None, help="A/B: 1 / 0 residual through [R | I] k-blocks / per-lane loads")
    ap.add_argument("--tma-store", type=int, default=None, help="A/B: 1 / 0 lean epilogues store through TMA / per lane")
    ap.add_argument("--lean-epilogue", type=int, default=None, help="A/B: 1 / 0 specialised / general GEMM epilogue code")
    ap.add_argument("--temporal-rows", type=int, default=None, help="A/B: 1 / 0 row-coalesced / per-head temporal attention kernel")
    ap.add_argument("--fuse-ln", type=int, default=None, help="A/B: 1 / 0 LayerNorm in the GEMM epilogue / as its own pass")
    ap.add_argument("--no-tc", action="store_true", help="debug: CUDA-core kernels only")
    ap.add_argument("--no-graph", action="store_true", help="debug: eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops-out", default=None, help="write the per-operator event-time table of one step to this file")
    ap.add_argument("--cpu-budget", type=float, default=150.0)
    ap.add_argument("--ncu-step", action="store_true", help="profiling aid: one eager step between cudaProfilerStart/Stop, then exit")
    ap.add_argument("--frame-shards", type=int, default=int(os.environ.get("MMGT_FRAME_SHARDS", "0")),
                    help="ranks that split the frames of one context window (SURVEY 8e level 3); must divide --gpus; "
                         "0 = auto: whole forwards, plus frame-sharded leftovers when they do not divide over the ranks")
    ap.add_argument("--shard-remainder", action="store_true",
                    help="with --frame-shards k: frame-shard only the forwards left over by the whole deal")
    ap.add_argument("--pdl", type=int, default=None, help="1 / 0: force programmatic dependent launch on / off (default: library default)")
    ap.add_argument("--unfused-exchange", action="store_true",
                    help="A/B: GEMM + stand-alone row-exchange copy instead of peer stores from the GEMM epilogue")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the mmgt_b200 path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
