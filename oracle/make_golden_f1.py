"""Golden fixtures for the one-shot conditioning modules.  TEST INFRASTRUCTURE.

Runs ONLY where /root/reference exists: imports the reference's own PoseGuider and AudioProjModel unchanged (on the
stand-in ``diffusers`` of oracle/ref_shim, which they use for ``ModelMixin`` only), loads deterministic synthetic weights
(oracle/weights.py; the zero-initialised conv_out gets N(0, 0.02^2) so it is numerically visible), runs them on small seeded
inputs and writes tests/golden/conditioning.npz (inputs are regenerated from the seeds; outputs and state-dict specs stored).

Usage:  python -m oracle.make_golden_f1
"""
import importlib
import json
import os

import numpy as np
import torch

from . import reference_loader as RL
from .conditioning import audio_proj_forward, pose_guider_forward
from .weights import make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

POSE_CFG = dict(conditioning_embedding_channels=320, block_out_channels=(16, 32, 64, 128))     # scripts/pose2vid.py:160-163
AUDIO_CFG = dict(seq_len=5, blocks=12, channels=768, intermediate_dim=512, output_dim=768, context_tokens=32)


def pose_input(seed=5, frames=3, size=64):
    return torch.rand(1, 3, frames, size, size, generator=torch.Generator().manual_seed(seed))


def audio_input(seed=6, frames=2, cfg=AUDIO_CFG):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, frames, cfg["seq_len"], cfg["blocks"], cfg["channels"], generator=g)


def zero_init_visible(sd, key_prefix, seed):
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.startswith(key_prefix):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    return sd


def main():
    RL._activate()
    pg_mod = importlib.import_module("src.models.pose_guider")
    ap_mod = importlib.import_module("src.models.audio_proj")
    out = {}
    specs = {}
    with torch.no_grad():
        pg = pg_mod.PoseGuider(**POSE_CFG)
        spec = [(k, tuple(v.shape)) for k, v in pg.state_dict().items()]
        sd = zero_init_visible(make_state_dict(spec, seed=3), "conv_out", seed=31)
        pg.load_state_dict(sd, strict=True)
        x = pose_input()
        ref = pg(x)
        mine = pose_guider_forward(sd, x, POSE_CFG["block_out_channels"])
        err = float((mine - ref).norm() / ref.norm())
        print(f"PoseGuider: oracle vs reference rel-L2 {err:.3e}, out {tuple(ref.shape)}")
        assert err < 1e-6
        out["pose_out"] = ref.numpy()
        specs["pose_guider"] = [[k, list(s)] for k, s in spec]

        ap = ap_mod.AudioProjModel(**AUDIO_CFG)
        spec = [(k, tuple(v.shape)) for k, v in ap.state_dict().items()]
        sd = make_state_dict(spec, seed=4)
        ap.load_state_dict(sd, strict=True)
        a = audio_input()
        ref = ap(a)
        mine = audio_proj_forward(sd, a, AUDIO_CFG["context_tokens"], AUDIO_CFG["output_dim"])
        err = float((mine - ref).norm() / ref.norm())
        print(f"AudioProjModel: oracle vs reference rel-L2 {err:.3e}, out {tuple(ref.shape)}")
        assert err < 1e-6
        out["audio_out"] = ref.numpy().astype(np.float32)
        specs["audio_proj"] = [[k, list(s)] for k, s in spec]
    np.savez_compressed(os.path.join(GOLD, "conditioning.npz"), **out)
    with open(os.path.join(GOLD, "conditioning_spec.json"), "w") as f:
        json.dump(specs, f)
    print("wrote", os.path.join(GOLD, "conditioning.npz"))


if __name__ == "__main__":
    main()
