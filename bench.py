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


def synthetic_video(L, latent, pinned=True):
    """Host-side (pinned) whole-video inputs in the layout Pose2VideoPipeline holds before the loop."""
    g = torch.Generator().manual_seed(42)
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
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl
    from mmgt_b200.pipeline_pose2vid_long import DenoiseLoop
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
    L = args.frames
    t0 = time.time()
    unet = build_unet(dev, cdt)
    host = synthetic_video(L, LATENT)
    log(f"[rank {rank}] model + inputs built in {time.time() - t0:.1f}s")
    ctl = ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
    ctl.set_banks([b.to(dev) for b in host["banks"]])
    sched = DDIMSchedule.from_config()

    # Schedule.  Default ("auto"): whole windows (B=2), then single-branch forwards, dealt evenly to the ranks; forwards that
    # still do not divide (10 windows on 8 GPUs: one window each, 4 forwards left) are shared by pairs of ranks as frame
    # shards (2.5 forwards of work per rank instead of 3 / 2).
    shards, remainder = args.frame_shards, args.shard_remainder
    if shards == 0:
        from mmgt_b200.context import get_context_scheduler
        from mmgt_b200.pipeline_pose2vid_long import plan_units_mixed
        n_windows = len(list(get_context_scheduler("uniform")(0, N_STEPS, L, 12, 1, 4)))
        shards, remainder = 1, False
        if world % 2 == 0 and any(plan_units_mixed(n_windows, 2, world, 2)[1]):
            shards, remainder = 2, True
    if world % shards:
        raise SystemExit(f"bench.py: --frame-shards {shards} must divide the number of ranks {world}")
    shared = {"shards": shards, "remainder": remainder}

    def make_loop(d):
        for attempt in range(2):
            loop = DenoiseLoop(unet, sched, N_STEPS, GUIDANCE, motion_scale=[1.0, 1.0, 2.0], rank=rank, world_size=world,
                               frame_shards=shared["shards"], shard_group=shared.get("group"),
                               shard_remainder=shared["remainder"])
            try:
                loop.prepare(d["latents"], d["pose"], d["audio"], d["full"], d["face"], d["lip"], d["ehs"])
            except RuntimeError as e:
                if attempt == 0 and args.frame_shards == 0 and "peer memory" in str(e):   # raised on every rank together
                    log(f"[rank {rank}] {e}; falling back to whole forwards only")
                    shared.update(shards=1, remainder=False)
                    continue
                raise
            shared["group"] = loop.shard_group        # one set of peer buffers serves every loop of this process
            return loop

    torch.cuda.synchronize()
    t_first0 = time.perf_counter()
    d = to_device(host, dev)
    loop = make_loop(d)
    eng = loop.eng
    if args.no_tc:
        eng.ctx.set_tensor_cores(False)
    eng.unfused_exchange = args.unfused_exchange
    if args.pdl is not None:
        eng.ctx.set_pdl(bool(args.pdl))
    if not args.no_graph:
        t0 = time.time()
        loop.capture_graph()
        log(f"[rank {rank}] CUDA graph of one step captured in {time.time() - t0:.1f}s ({loop.graph_launches} kernel launches)")
    torch.cuda.synchronize()
    first_video_prepare_s = time.perf_counter() - t_first0      # one-off per process and video shape (includes graph capture)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up + K timed steps, CUDA events, max over ranks
    for i in range(args.warmup):
        loop.step(i % N_STEPS)
    barrier()
    if args.ncu_step:
        # `ncu --profile-from-start off --metrics gpu__time_duration.sum ... python bench.py --ncu-step`: exactly one DDIM
        # step (eager launches) between cudaProfilerStart / Stop -> the launch list under profiles/; prints no bench line.
        graph, loop._graph = loop._graph, None
        torch.cuda.cudart().cudaProfilerStart()
        loop.step(args.warmup % N_STEPS)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        loop._graph = graph
        log("[ncu-step] one step profiled; no bench line is printed under a profiler")
        return
    sampler = ClockSampler(local)
    sampler.start()
    n0 = eng.ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loop.step((args.warmup + i) % N_STEPS)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.ctx.launches() - n0 + (loop.graph_launches * args.steps if loop._graph is not None else 0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    value = L / (N_STEPS * ms_per_step / 1e3)

    # ---- end-to-end through the public loop API with HOST buffers: a NEW video of the same shape.  Its conditioning is
    #      uploaded from pinned host memory and written into the loop's static buffers (DenoiseLoop.reload: the CUDA graph
    #      captured for the first video keeps serving) -- timed and amortised over the 30 steps; every step then does the
    #      H2D copy of the latents and the D2H read of the updated latents.
    barrier()
    t_prep0 = time.perf_counter()
    d2 = to_device(host, dev)
    loop2 = loop.reload(d2["latents"], d2["pose"], d2["audio"], d2["full"], d2["face"], d2["lip"], d2["ehs"])
    torch.cuda.synchronize()
    prep_s = time.perf_counter() - t_prep0
    lat_host = host["latents"]
    out_host = torch.empty_like(lat_host).pin_memory()
    e_steps = max(1, min(args.steps, 3))
    barrier()
    t_e0 = time.perf_counter()
    for i in range(e_steps):
        loop2.latents.copy_(lat_host, non_blocking=True)
        loop2.step(i % N_STEPS)
        out_host.copy_(loop2.latents, non_blocking=True)
        torch.cuda.synchronize()
    e2e_step = (time.perf_counter() - t_e0) / e_steps
    e2e_t = torch.tensor([e2e_step + prep_s / N_STEPS], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = L / (N_STEPS * float(e2e_t))
    lat_bytes = lat_host.numel() * 4
    del d2

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launch stream over one more step
    roof = None
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
        whole = UNITS_PER_STEP * (L / 80.0) * FLOP_PER_UNIT / (ms_per_step * 1e-3) / 1e12 / world
        roof["whole_step_tflops_per_gpu"] = whole
        roof["whole_step_frac_of_peak"] = whole / pk["tf_sustained"]

    if shared.get("group") is not None:
        shared["group"].check()                   # raises if any peer barrier timed out during the run
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(unet, steps=1, warmup=0, budget_s=args.cpu_budget)

    if rank == 0:
        line = dict(metric="UNet3D denoise frames/s @512x512 80f (30-step CFG DDIM, 10 windows x 12 frames)", value=value,
                    unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=dict(workload=f"pose2vid 512x512 (64x64 latent), {L} frames, 30 DDIM steps, CFG 3.5, full-width "
                                         "UNet3D random-init; step = 1 DDIM step = 10 windows x 2 CFG branches",
                                parallelism=_describe(world, shared["shards"], shared["remainder"], loop),
                                l2_policy="per-step working set (weights 2.8 GB + activations) exceeds the 126 MB L2",
                                tensor_cores=not args.no_tc, programmatic_dependent_launch=eng.ctx.pdl()),
                    clocks=clocks, gpu_launches=launches,
                    e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=lat_bytes + h2d_bytes(host) // N_STEPS,
                             d2h_bytes_per_step=lat_bytes, prepare_s=prep_s, step_s=e2e_step,
                             first_video_prepare_s=first_video_prepare_s),
                    roofline=roof, cpu_baseline=cpu, frame_evals_per_s=UNITS_PER_STEP * 12 * (L / 80.0) / (ms_per_step / 1e3))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_reference(unet_or_none, steps, warmup, budget_s):
    """Times the oracle (CPU restatement of the reference algorithm, plain PyTorch fp32) on the host cores.
    One sample = one (window, CFG-branch) forward of config 2: B=1, F=12, 64x64 latent; 600 of them make a video."""
    from oracle.unet3d import UNetSpec, unet3d_forward, spatial_block_prefixes, spatial_block_width
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
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            unet3d_forward(sd, spec, sample, 500, ehs, aud, pose, masks[0], masks[1], masks[2], [1.0, 1.0, 2.0], banks,
                           ref_index=[1], apply_motion_scale=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            log(f"[cpu] unit forward {i}: {dt:.1f}s")
            if time.perf_counter() - t_start > budget_s and times:
                break
    t_unit = sum(times) / len(times)
    units_per_video = N_STEPS * UNITS_PER_STEP
    return dict(value=VIDEO_LENGTH / (units_per_video * t_unit), unit="frames/s", cores=cores, kind="port",
                sample=f"{len(times)} x one (window, CFG-branch) UNet3D forward of the same workload (B=1, 12 frames, 64x64 "
                       f"latent, fp32) = 1/{units_per_video} of the 30-step video, extrapolated; {t_unit:.1f}s each",
                seconds_per_unit=t_unit, steps_done=len(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_reference(None, steps=args.steps, warmup=min(args.warmup, 1), budget_s=args.cpu_budget)
    t_step = cpu["seconds_per_unit"] * UNITS_PER_STEP
    line = dict(impl="reference", metric="UNet3D denoise frames/s @512x512 80f (30-step CFG DDIM, 10 windows x 12 frames)",
                value=cpu["value"], unit="frames/s", n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=cpu["steps_done"],
                warmup=min(args.warmup, 1), ms_per_step=t_step * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload="pose2vid 512x512 (64x64 latent), 80 frames, 30 DDIM steps, CFG 3.5, full-width UNet3D "
                                     "random-init; CPU arm times a bounded sample (one window x one CFG branch) and extrapolates"),
                cpu_baseline=cpu, e2e=dict(value=cpu["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--frames", type=int, default=VIDEO_LENGTH)
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
