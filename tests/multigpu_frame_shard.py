"""torchrun --nproc-per-node 2 tests/multigpu_frame_shard.py  (one process per GPU, NCCL + CUDA IPC peer memory)

Two DDIM steps of a 20-frame video (3 overlapping windows, CFG) with every window's frames split over the two
GPUs (DenoiseLoop(frame_shards=2)): the motion modules exchange rows through peer stores from the GEMM epilogue.
Checked on every rank against the oracle's denoise loop (north-star tolerances) and against the window-parallel
schedule (frame_shards=1) of the same two ranks.  Prints FRAME-SHARD PARITY OK on rank 0."""
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
    L, latent, n_steps = 20, 16, 30
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
        for shards in (world, 1):
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
        diff = rel_l2(res[world], res[1])
        print(f"[rank {rank}] {dtype}: frame-sharded vs window-parallel rel-L2 {diff:.3e}", flush=True)
        ok = ok and diff < (1e-5 if dtype == torch.float32 else 5e-3)
        # Mixed schedule (what bench.py uses when the forwards do not divide over the ranks): without CFG the 3 windows
        # are 3 forwards for 2 ranks -> one whole forward each + one shared as two frame shards; vs all-whole dealing.
        attach_banks(unet, spec, banks, cfg=False)
        nocfg = dict(audio=d["audio"][1:2], ehs=d["encoder_hidden_states"][1:2],
                     masks=[[m[L:] for m in d[name]] for name in ("full_mask", "face_mask", "lip_mask")])
        mixed = {}
        for remainder in (True, False):
            loop = DenoiseLoop(unet, DDIMSchedule.from_config(), n_steps, 1.0, motion_scale=inp["motion_scale"], rank=rank,
                               world_size=world, frame_shards=world if remainder else 1, shard_remainder=remainder)
            loop.prepare(d["latents"], d["pose_fea"], nocfg["audio"], nocfg["masks"][0], nocfg["masks"][1], nocfg["masks"][2],
                         nocfg["ehs"])
            if remainder:
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
