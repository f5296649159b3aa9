"""Pin the oracle to the reference and write the golden fixtures.  TEST INFRASTRUCTURE.

Runs ONLY where /root/reference exists (this container).  Imports the reference's modules
unchanged (``reference_loader``), loads the deterministic synthetic weights, runs
``UNet3DConditionModel.forward`` with ``ReferenceAttentionControl`` in read mode, compares with
``oracle.unet3d`` and stores small fixtures under tests/golden/:

  statedict_spec.json      : the reference's 1526 state-dict keys + shapes (full + tiny width)
  unet_tiny_*.npz          : tiny-width UNet, outputs of both code branches + per-module taps
  unet_full_cfg1.npz       : full-width UNet at BASELINE config 1 (32x32 latent, 8 frames, CFG)
  windows.json, ddim.json  : context windows and DDIM tables

Usage:  python -m oracle.make_golden [--skip-full]
"""
import argparse
import json
import os
import time

import numpy as np
import torch

from . import reference_loader as RL
from .sampler import DDIM, uniform_windows
from .synthetic import make_banks, make_inputs, window_inputs
from .unet3d import UNetSpec, bank_pairing_order, unet3d_forward
from .weights import make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_reference(unet, mods, win, banks, spec, timestep, branch, cfg=True):
    """branch: 'scripts' = train()+gradient checkpointing (what scripts/pose2vid.py runs, fact 4)
    or 'eval'."""
    msa = mods["mutual_self_attention"]
    attn_mod = mods["attention"]
    reader = msa.ReferenceAttentionControl(unet, do_classifier_free_guidance=cfg, mode="read",
                                           batch_size=1, fusion_blocks="full")
    blocks = [m for m in msa.torch_dfs(unet) if isinstance(m, attn_mod.TemporalBasicTransformerBlock)]
    blocks = sorted(blocks, key=lambda m: -m.norm1.normalized_shape[0])
    order = bank_pairing_order(spec)
    assert len(blocks) == len(order) == 16
    name_of = {id(m): n for n, m in unet.named_modules()}
    for blk, pre in zip(blocks, order):
        assert name_of[id(blk)] == pre + ".transformer_blocks.0", (name_of[id(blk)], pre)
        blk.bank = [banks[pre].clone() if cfg else banks[pre][0:1].clone()]
    if branch == "scripts":
        unet.train()
        unet.enable_gradient_checkpointing()
    else:
        unet.eval()
    taps = {}
    hooks = []
    for n, m in unet.named_modules():
        cls = type(m).__name__
        if cls in ("ResnetBlock3D", "Transformer3DModel", "VanillaTemporalModule", "Downsample3D", "Upsample3D"):
            def hook(mod, args, out, n=n):
                o = out[0] if isinstance(out, tuple) else (out.sample if hasattr(out, "sample") else out)
                taps[n] = o.detach().clone()
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = unet(win["sample"], torch.tensor(timestep), encoder_hidden_states=win["encoder_hidden_states"],
                       audio_embedding=win["audio_embedding"], pose_cond_fea=win["pose_cond_fea"],
                       full_mask=win["full_mask"], face_mask=win["face_mask"], body_mask=win["body_mask"],
                       motion_scale=win["motion_scale"], return_dict=False)[0]
    for h in hooks:
        h.remove()
    return out, taps


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden_case(tag, boc, latent, frames, timestep, save_taps):
    spec = UNetSpec(block_out_channels=tuple(boc))
    unet, mods = RL.build_reference_unet(block_out_channels=boc)
    shapes = [(k, tuple(v.shape)) for k, v in unet.state_dict().items()]
    sd = make_state_dict(shapes, seed=0)
    unet.load_state_dict(sd, strict=True)
    inp = make_inputs(spec, frames, latent)
    banks = make_banks(spec, latent)
    win = window_inputs(inp, list(range(frames)))
    res = {}
    for branch in ("scripts", "eval"):
        t0 = time.time()
        ref, ref_taps = run_reference(unet, mods, win, banks, spec, timestep, branch)
        t_ref = time.time() - t0
        taps = {}
        t0 = time.time()
        with torch.no_grad():
            mine = unet3d_forward(sd, spec, win["sample"], timestep, win["encoder_hidden_states"],
                                  win["audio_embedding"], win["pose_cond_fea"], win["full_mask"],
                                  win["face_mask"], win["body_mask"], win["motion_scale"], banks,
                                  ref_index=[None, 1], apply_motion_scale=(branch == "scripts"), taps=taps)
        t_or = time.time() - t0
        err = rel_l2(mine, ref)
        worst = max(((rel_l2(taps[k], ref_taps[k]), k) for k in taps if k in ref_taps), default=(0, ""))
        print(f"[{tag}/{branch}] oracle-vs-reference rel-L2 {err:.3e}  worst tap {worst[0]:.3e} @ {worst[1]}"
              f"  (ref {t_ref:.1f}s oracle {t_or:.1f}s) |out| {float(ref.norm()):.3f}")
        assert err < 2e-5, "oracle does not reproduce the reference"
        res[branch] = (ref, ref_taps)
    # no-CFG reader (B=1): every frame attends to [self || bank[0]]
    win1 = {k: (v[1:2] if torch.is_tensor(v) and v.shape[0] == 2 else v) for k, v in win.items()}
    for k in ("full_mask", "face_mask", "body_mask"):
        win1[k] = [m[frames:] for m in win[k]]
    ref1, _ = run_reference(unet, mods, win1, banks, spec, timestep, "scripts", cfg=False)
    with torch.no_grad():
        mine1 = unet3d_forward(sd, spec, win1["sample"], timestep, win1["encoder_hidden_states"],
                               win1["audio_embedding"], win1["pose_cond_fea"], win1["full_mask"],
                               win1["face_mask"], win1["body_mask"], win1["motion_scale"], banks,
                               ref_index=[0], apply_motion_scale=True)
    e1 = rel_l2(mine1, ref1)
    print(f"[{tag}/nocfg] oracle-vs-reference rel-L2 {e1:.3e}")
    assert e1 < 2e-5
    out = dict(out_scripts=res["scripts"][0].numpy(), out_eval=res["eval"][0].numpy(), out_nocfg=ref1.numpy(),
               timestep=np.int64(timestep), latent=np.int64(latent), frames=np.int64(frames),
               block_out_channels=np.array(boc))
    if save_taps:
        for k, v in res["scripts"][1].items():
            out["tap:" + k] = v.numpy().astype(np.float16) if v.numel() > 70000 else v.numpy()
    else:
        for k, v in res["scripts"][1].items():   # statistics only (fixture size)
            out["tapstat:" + k] = np.array([float(v.mean()), float(v.std()), float(v.abs().max())], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, f"unet_{tag}.npz"), **out)
    return shapes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    spec_json = {}
    spec_json["tiny"] = golden_case("tiny", [64, 128, 256, 256], latent=16, frames=4, timestep=500, save_taps=False)
    if not args.skip_full:
        spec_json["full"] = golden_case("full_cfg1", [320, 640, 1280, 1280], latent=32, frames=8, timestep=500,
                                        save_taps=False)
    else:
        unet, _ = RL.build_reference_unet()
        spec_json["full"] = [(k, tuple(v.shape)) for k, v in unet.state_dict().items()]
    with open(os.path.join(GOLD, "statedict_spec.json"), "w") as f:
        json.dump({k: [[n, list(s)] for n, s in v] for k, v in spec_json.items()}, f)
    with open(os.path.join(GOLD, "windows.json"), "w") as f:
        ref_uniform = RL.load_reference_modules()["context"].uniform
        cases = {}
        for L in (8, 12, 13, 24, 80, 160):
            w = [list(map(int, x)) for x in ref_uniform(0, 30, L, 12, 1, 4)]
            assert w == uniform_windows(0, L, 12, 1, 4)
            cases[str(L)] = w
        json.dump(cases, f)
    d = DDIM()
    with open(os.path.join(GOLD, "ddim.json"), "w") as f:
        json.dump(dict(timesteps30=d.timesteps(30), alphas_cumprod_sample={str(t): float(d.alphas_cumprod[t])
                                                                          for t in (0, 32, 500, 966, 999)}), f)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
