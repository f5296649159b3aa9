"""Golden fixture at the BENCHED shapes (BASELINE config 2): one DDIM step of the 80-frame / 10-window video.
TEST INFRASTRUCTURE; runs only where /root/reference exists (this container), ~10 min on 8 cores.

The reference's own ``UNet3DConditionModel`` (imported unchanged through ``reference_loader``; full width, 1.40 G
parameters, deterministic synthetic weights, ``train()`` + gradient checkpointing = the scripts' branch) is called once
per context window exactly as ``Pose2VideoPipeline.__call__`` calls it (pipeline_pose2vid_long.py:554-620, restated in
``oracle.sampler.denoise_step``), with CFG, non-zero audio on the cond branch and three distinct motion masks
(``oracle.synthetic.make_inputs``) -- which also makes it the BASELINE config-4 content (MM-HAA live).  Stored:

  tests/golden/config2_step.npz   latents_out (1,4,80,64,64) after the step at timestep index 15 (t = 499),
                                  (the CFG-combined prediction v follows from it: latents_out = cx * latents + cv * v),
                                  pred_w0 (2,4,12,64,64) the raw UNet output of window 0 (the config-2 window shape)

Inputs are NOT stored: ``make_inputs(spec, 80, 64)`` / ``make_banks`` / ``make_state_dict`` regenerate them bit-identically.

Usage:  python -m oracle.make_golden_config2
"""
import os
import time

import numpy as np
import torch

from . import reference_loader as RL
from .make_golden import GOLD, rel_l2, run_reference
from .sampler import DDIM, denoise_step, uniform_windows
from .synthetic import make_banks, make_inputs
from .unet3d import UNetSpec, unet3d_forward
from .weights import make_state_dict

STEP_INDEX, N_STEPS, L, LATENT, GUIDANCE = 15, 30, 80, 64, 3.5


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    boc = [320, 640, 1280, 1280]
    spec = UNetSpec(block_out_channels=tuple(boc))
    unet, mods = RL.build_reference_unet(block_out_channels=boc)
    shapes = [(k, tuple(v.shape)) for k, v in unet.state_dict().items()]
    sd = make_state_dict(shapes, seed=0)
    unet.load_state_dict(sd, strict=True)
    inp = make_inputs(spec, L, LATENT)
    banks = make_banks(spec, LATENT)
    windows = uniform_windows(0, L)
    assert len(windows) == 10
    ddim = DDIM()
    t = ddim.timesteps(N_STEPS)[STEP_INDEX]
    preds = []

    def unet_fn(sample, tt, ehs, aud, pose, full, face, lip, ms):
        t0 = time.time()
        win = dict(sample=sample, encoder_hidden_states=ehs, audio_embedding=aud, pose_cond_fea=pose, full_mask=full,
                   face_mask=face, body_mask=lip, motion_scale=ms)
        out, _ = run_reference(unet, mods, win, banks, spec, tt, "scripts")
        preds.append(out)
        print(f"  reference window forward {len(preds)}/10: {time.time() - t0:.1f}s", flush=True)
        return out
    lat, v = denoise_step(unet_fn, inp["latents"].clone(), t, N_STEPS, ddim, GUIDANCE, windows, inp["pose_fea"], inp["audio"],
                          inp["full_mask"], inp["face_mask"], inp["lip_mask"], inp["encoder_hidden_states"],
                          inp["motion_scale"])
    # the oracle port against the reference at this shape (window 0)
    c = windows[0]
    g = lambda ms: [m.view(2, L, -1)[:, c, :].reshape(-1, m.shape[-1]) for m in ms]  # noqa: E731
    with torch.no_grad():
        mine = unet3d_forward(sd, spec, inp["latents"][:, :, c].repeat(2, 1, 1, 1, 1), t, inp["encoder_hidden_states"],
                              inp["audio"][:, c], inp["pose_fea"][:, :, c].repeat(2, 1, 1, 1, 1), g(inp["full_mask"]),
                              g(inp["face_mask"]), g(inp["lip_mask"]), inp["motion_scale"], banks, ref_index=[None, 1],
                              apply_motion_scale=True)
    err = rel_l2(mine, preds[0])
    print(f"oracle port vs reference at the config-2 window shape: rel-L2 {err:.3e}")
    assert err < 2e-5
    np.savez_compressed(os.path.join(GOLD, "config2_step.npz"), latents_out=lat.numpy(), pred_w0=preds[0].numpy(),
                        timestep=np.int64(t), step_index=np.int64(STEP_INDEX), n_steps=np.int64(N_STEPS),
                        oracle_vs_reference_w0=np.float64(err))
    print("written", os.path.join(GOLD, "config2_step.npz"))


if __name__ == "__main__":
    main()
