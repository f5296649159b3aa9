"""Shared test helpers (tests may import oracle/; the product package may not)."""
import json
import os

import numpy as np
import torch

from oracle.reference_loader import SD15_CFG, UNET_ADDITIONAL_KWARGS
from oracle.synthetic import make_banks, make_inputs, window_inputs
from oracle.unet3d import UNetSpec, bank_pairing_order
from oracle.weights import make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = (64, 128, 256, 256)
FULL = (320, 640, 1280, 1280)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def statedict_spec(tag):
    with open(os.path.join(GOLD, "statedict_spec.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)[tag]]


def synthetic_state_dict(tag, seed=0):
    return make_state_dict(statedict_spec(tag), seed=seed)


def build_cuda_unet(boc, sd, device="cuda", param_dtype=torch.float32, compute_dtype=None):
    from mmgt_b200.unet_3d import UNet3DConditionModel
    cfg = dict(SD15_CFG)
    cfg["block_out_channels"] = list(boc)
    unet = UNet3DConditionModel.from_config(cfg, **UNET_ADDITIONAL_KWARGS)
    unet.load_state_dict(sd, strict=True)
    unet.to(device=device, dtype=param_dtype)
    if compute_dtype is not None:
        unet.set_compute_dtype(compute_dtype)
    return unet


def attach_banks(unet, spec: UNetSpec, banks, cfg=True):
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl
    ctl = ReferenceAttentionControl(unet, do_classifier_free_guidance=cfg, mode="read", batch_size=1, fusion_blocks="full")
    order = bank_pairing_order(spec)
    ctl.set_banks([banks[p] if cfg else banks[p][0:1] for p in order])
    return ctl


def to_dev(win, device):
    out = {}
    for k, v in win.items():
        if torch.is_tensor(v):
            out[k] = v.to(device)
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [t.to(device) for t in v]
        else:
            out[k] = v
    return out


def run_cuda_unet(unet, win, timestep, device="cuda"):
    w = to_dev(win, device)
    return unet(w["sample"], torch.tensor(timestep), encoder_hidden_states=w["encoder_hidden_states"],
                audio_embedding=w["audio_embedding"], pose_cond_fea=w["pose_cond_fea"], full_mask=w["full_mask"],
                face_mask=w["face_mask"], body_mask=w["body_mask"], motion_scale=w["motion_scale"], return_dict=False)[0]
