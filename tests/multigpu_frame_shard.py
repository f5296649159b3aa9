"""torchrun --nproc-per-node N tests/multigpu_frame_shard.py  (N = 2, 4 or 8: one process per GPU, NCCL + CUDA IPC peer memory)

Two DDIM steps of a short video (CFG, overlapping windows) under every schedule DenoiseLoop / bench.py can pick at N ranks:
every window as k frame shards (k = 2, 4 where it divides: the motion modules exchange rows through peer stores from
the GEMM epilogue), whole forwards dealt over the ranks, and the mixed "whole + leftovers shared by pairs" schedule
(bench.py's default at 4 and 8 GPUs).  Checked on every rank against the oracle's denoise loop (north-star tolerances)
and against each other.  Prints FRAME-SHARD PARITY OK on rank 0.  Collected by pytest through
tests/test_frame_shard_gpu.py::test_multi_gpu_schedules_match_oracle."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from helpers import TINY, attach_banks, build_cuda_unet, rel_l2, synthetic_state_dict, to_dev  # noqa: E402
from oracle.sampler import DDIM, denoise_step, uniform_windows  # noqa: E402
from oracle.synthetic import make_banks, make_inputs  # noqa: E402
from oracle.unet3d import UNetSpec, unet3d_forward  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mmgt_b200.pipeline_pose2vid_long import DenoiseLoop
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    spec = UNetSpec(block_out_channels=TINY)
    sd = synthetic_state_dict("tiny")
    # 20 frames = 3 windows (6 forwards); 44 frames = 6 windows: on 8 ranks 12 forwards = 1 each + 4 shared by the 4 pairs
    L, latent, n_steps = (20 if world <= 4 else 44), 16, 30
    inp = make_inputs(spec, L, latent, seed=11)
    banks = make_banks(spec, latent)
    windows = uniform_windows(0, L)

    def unet_fn(sample, t, ehs, aud, pose, full, face, lip, ms):
        with torch.no_grad():
            return unet3d_forward(sd, spec, sample, t, ehs, aud, pose, full, face, lip, ms, banks, ref_index=[None, 1],
                                  apply_motion_scale=True)
    ddim = DDIM()
    lat_ref = inp["latents"].clone()
    for t in ddim.timesteps(n_steps)[:2]:
        lat_ref, _ = denoise_step(unet_fn, lat_ref, t, n_steps, ddim, 3.5, windows, inp["pose_fea"], inp["audio"],
                                  inp["full_mask"], inp["face_mask"], inp["lip_mask"], inp["encoder_hidden_states"],
                                  inp["motion_scale"])
    ok = True
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 1e-2)):
        unet = build_cuda_unet(TINY, sd, device=dev, compute_dtype=dtype)
        unet.train()
        unet.enable_gradient_checkpointing()
        attach_banks(unet, spec, banks, cfg=True)
        d = to_dev(inp, dev)
        res = {}
        shard_opts = [k for k in (2, 4) if world % k == 0] + [1]
        for shards in shard_opts:
            loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 3.5, motion_scale=inp["motion_scale"], rank=rank,
                               world_size=world, frame_shards=shards)
            loop.prepare(d["latents"], d["pose_fea"], d["audio"], d["full_mask"], d["face_mask"], d["lip_mask"],
                         d["encoder_hidden_states"])
            if dtype == torch.bfloat16:
                loop.capture_graph()          # the peer stores and the flag barrier replay from a CUDA graph
            loop.step(0)
            lat = loop.step(1).clone()
            if loop.shard_group is not None:
                loop.shard_group.check()
            res[shards] = lat
            err = rel_l2(lat, lat_ref)
            print(f"[rank {rank}] {dtype} frame_shards={shards}: latents rel-L2 vs oracle {err:.3e} (tol {tol})", flush=True)
            ok = ok and err < tol
            if loop.shard_group is not None:
                loop.shard_group.close()
        for k in shard_opts[:-1]:
            diff = rel_l2(res[k], res[1])
            print(f"[rank {rank}] {dtype}: {k} frame shards vs window-parallel rel-L2 {diff:.3e}", flush=True)
            ok = ok and diff < (1e-5 if dtype == torch.float32 else 5e-3)
        # the CFG schedule bench.py picks by default at this world size (plan_units_mixed: whole forwards + leftovers
        # shared by pairs), through the CUDA graph in bf16
        loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 3.5, motion_scale=inp["motion_scale"], rank=rank,
                           world_size=world, frame_shards=2, shard_remainder=True)
        loop.prepare(d["latents"], d["pose_fea"], d["audio"], d["full_mask"], d["face_mask"], d["lip_mask"],
                     d["encoder_hidden_states"])
        if dtype == torch.bfloat16:
            loop.capture_graph()
        loop.step(0)
        lat = loop.step(1).clone()
        n_shared = sum(1 for _, _, sh in loop.units if sh)
        err = rel_l2(lat, lat_ref)
        print(f"[rank {rank}] {dtype} default mixed schedule ({len(loop.units) - n_shared} whole + {n_shared} shared units on "
              f"this rank): latents rel-L2 vs oracle {err:.3e} (tol {tol})", flush=True)
        ok = ok and err < tol
        if loop.shard_group is not None:
            loop.shard_group.check()
        loop.close()
        # Mixed schedule (what bench.py uses when the forwards do not divide over the ranks): without CFG the 3 windows
        # are 3 forwards for 2 ranks -> one whole forward each + one shared as two frame shards; vs all-whole dealing.
        attach_banks(unet, spec, banks, cfg=False)
        nocfg = dict(audio=d["audio"][1:2], ehs=d["encoder_hidden_states"][1:2],
                     masks=[[m[L:] for m in d[name]] for name in ("full_mask", "face_mask", "lip_mask")])
        mixed = {}
        for remainder in (True, False):
            loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 1.0, motion_scale=inp["motion_scale"], rank=rank,
                               world_size=world, frame_shards=2 if remainder else 1, shard_remainder=remainder)
            loop.prepare(d["latents"], d["pose_fea"], nocfg["audio"], nocfg["masks"][0], nocfg["masks"][1], nocfg["masks"][2],
                         nocfg["ehs"])
            if remainder and world == 2:
                assert sorted(sh for _, _, sh in loop.units) == [False, True], loop.units
            if dtype == torch.bfloat16:
                loop.capture_graph()
            loop.step(0)
            mixed[remainder] = loop.step(1).clone()
            if loop.shard_group is not None:
                loop.shard_group.check()
                loop.shard_group.close()
        diff = rel_l2(mixed[True], mixed[False])
        print(f"[rank {rank}] {dtype}: whole + shared-remainder schedule vs whole-only rel-L2 {diff:.3e}", flush=True)
        ok = ok and diff < (1e-5 if dtype == torch.float32 else 5e-3)
        del unet
        torch.cuda.empty_cache()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and float(flag) == 1.0:
        print("FRAME-SHARD PARITY OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
