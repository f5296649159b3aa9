"""CPU: the oracle restatement against the golden fixtures generated from the reference's own modules
(oracle/make_golden.py), plus the integer mask pyramid against live torchvision + Pillow."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, TINY, rel_l2, synthetic_state_dict
from oracle.mask_pyramid import preprocess_mov_mask, resize_bilinear_u8
from oracle.sampler import DDIM, uniform_windows
from oracle.synthetic import make_banks, make_inputs, synthetic_masks_u8, window_inputs
from oracle.unet3d import UNetSpec, bank_pairing_order, unet3d_forward


def _oracle_case(tag, boc, branch):
    g = np.load(os.path.join(GOLD, f"unet_{tag}.npz"))
    latent, frames, t = int(g["latent"]), int(g["frames"]), int(g["timestep"])
    spec = UNetSpec(block_out_channels=tuple(boc))
    sd = synthetic_state_dict("tiny" if tag == "tiny" else "full")
    inp = make_inputs(spec, frames, latent)
    banks = make_banks(spec, latent)
    win = window_inputs(inp, list(range(frames)))
    with torch.no_grad():
        if branch == "nocfg":
            w1 = {k: (v[1:2] if torch.is_tensor(v) and v.shape[0] == 2 else v) for k, v in win.items()}
            for k in ("full_mask", "face_mask", "body_mask"):
                w1[k] = [m[frames:] for m in win[k]]
            out = unet3d_forward(sd, spec, w1["sample"], t, w1["encoder_hidden_states"], w1["audio_embedding"],
                                 w1["pose_cond_fea"], w1["full_mask"], w1["face_mask"], w1["body_mask"], w1["motion_scale"],
                                 banks, ref_index=[0], apply_motion_scale=True)
        else:
            out = unet3d_forward(sd, spec, win["sample"], t, win["encoder_hidden_states"], win["audio_embedding"],
                                 win["pose_cond_fea"], win["full_mask"], win["face_mask"], win["body_mask"],
                                 win["motion_scale"], banks, ref_index=[None, 1], apply_motion_scale=(branch == "scripts"))
    return out, torch.from_numpy(g[f"out_{branch}"])


@pytest.mark.parametrize("branch", ["scripts", "eval", "nocfg"])
def test_oracle_matches_reference_golden_tiny(branch):
    out, gold = _oracle_case("tiny", TINY, branch)
    assert rel_l2(out, gold) < 2e-5


def test_scripts_and_eval_branches_differ():
    # fact 4: motion_scale only reaches MM-HAA in the scripts' branch -> the goldens must differ
    g = np.load(os.path.join(GOLD, "unet_tiny.npz"))
    assert rel_l2(torch.from_numpy(g["out_scripts"]), torch.from_numpy(g["out_eval"])) > 1e-4


def test_oracle_matches_reference_golden_full_width():
    out, gold = _oracle_case("full_cfg1", (320, 640, 1280, 1280), "scripts")
    assert rel_l2(out, gold) < 2e-5


def test_mask_pyramid_bit_exact_vs_pillow():
    from PIL import Image
    import torchvision.transforms as T
    face, lips = synthetic_masks_u8(6)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (4, 64, 64), dtype=np.uint8)
    for stack in (face, lips, noise):
        for img in stack:
            for s in (64, 32, 16, 8, 96, 48, 24, 12):
                ref = np.array(T.Resize((s, s))(Image.fromarray(img, "L")))
                assert np.array_equal(ref, resize_bilinear_u8(img, s, s)), s
    f_lvls, _ = preprocess_mov_mask(face, lips, 512)
    tt = T.Compose([T.Resize((32, 32)), T.ToTensor()])
    ref = torch.stack([tt(Image.fromarray(m, "L")) for m in face]).view(len(face), -1)
    assert torch.equal(ref, torch.from_numpy(f_lvls[1]))


def test_windows_and_ddim_match_golden():
    with open(os.path.join(GOLD, "windows.json")) as f:
        gold = json.load(f)
    for L, w in gold.items():
        assert uniform_windows(0, int(L), 12, 1, 4) == w
    assert len(gold["80"]) == 10 and len(gold["160"]) == 20
    with open(os.path.join(GOLD, "ddim.json")) as f:
        dd = json.load(f)
    d = DDIM()
    assert d.timesteps(30) == dd["timesteps30"]
    assert d.timesteps(30)[0] == 999 and d.timesteps(30)[-1] == 32
    for t, a in dd["alphas_cumprod_sample"].items():
        # float32 cumprod: the vectorised reduction order differs between host CPUs by an ulp or two
        assert abs(float(d.alphas_cumprod[int(t)]) - a) <= 2e-6 * max(abs(a), 1e-6)
    assert float(d.alphas_cumprod[999]) == 0.0


def test_bank_pairing_order_is_down_up_mid():
    order = bank_pairing_order(UNetSpec())
    assert len(order) == 16
    assert order[:6] == ["down_blocks.2.attentions.0", "down_blocks.2.attentions.1", "up_blocks.1.attentions.0",
                         "up_blocks.1.attentions.1", "up_blocks.1.attentions.2", "mid_block.attentions.0"]


def test_conditioning_oracle_matches_reference_golden():
    """SURVEY section 8f row f1 groundwork: the restated PoseGuider / AudioProjModel (oracle/conditioning.py) against outputs
    of the reference's own classes on the same seeded weights and inputs (oracle/make_golden_f1.py)."""
    import json
    from oracle.conditioning import audio_proj_forward, pose_guider_forward
    from oracle.make_golden_f1 import AUDIO_CFG, POSE_CFG, audio_input, pose_input, zero_init_visible
    from oracle.weights import make_state_dict
    g = np.load(os.path.join(GOLD, "conditioning.npz"))
    with open(os.path.join(GOLD, "conditioning_spec.json")) as f:
        specs = json.load(f)
    with torch.no_grad():
        sd = zero_init_visible(make_state_dict([(k, tuple(s)) for k, s in specs["pose_guider"]], seed=3), "conv_out", seed=31)
        out = pose_guider_forward(sd, pose_input(), POSE_CFG["block_out_channels"])
        assert out.shape == (1, 320, 3, 8, 8)                     # 64 x 64 pose image -> 8 x 8 x 320 features
        assert rel_l2(out, torch.from_numpy(g["pose_out"])) < 1e-5
        sd = make_state_dict([(k, tuple(s)) for k, s in specs["audio_proj"]], seed=4)
        out = audio_proj_forward(sd, audio_input(), AUDIO_CFG["context_tokens"], AUDIO_CFG["output_dim"])
        assert out.shape == (1, 2, 32, 768)
        assert rel_l2(out, torch.from_numpy(g["audio_out"])) < 1e-5


def test_mask_pyramid_on_the_reference_case_masks_bit_exact_vs_pillow():
    """The bundled real case (config/cases/oliver#103842_slice18_{face,lips}_mask.mp4 through the scripts' blur_mask
    front-end, oracle/make_golden_masks.py): the restated Pillow resize must reproduce live torchvision + Pillow bit for bit
    at every pyramid size of the 512^2 and 768^2 configurations, and ToTensor's u8 / 255."""
    from PIL import Image
    import torchvision.transforms as T
    g = np.load(os.path.join(GOLD, "real_masks.npz"))
    face, lips = g["face"], g["lips"]
    assert face.shape == lips.shape == (6, 64, 64) and face.dtype == np.uint8
    for stack in (face, lips):
        for img in stack:
            for s in (64, 32, 16, 8, 96, 48, 24, 12):
                ref = np.array(T.Resize((s, s))(Image.fromarray(img, "L")))
                assert np.array_equal(ref, resize_bilinear_u8(img, s, s)), s
    for image_size in (512, 768):
        f_lvls, l_lvls = preprocess_mov_mask(face, lips, image_size)
        for k in range(4):
            s = image_size // 8 >> k
            tt = T.Compose([T.Resize((s, s)), T.ToTensor()])
            for lv, stack in ((f_lvls, face), (l_lvls, lips)):
                ref = torch.stack([tt(Image.fromarray(m, "L")) for m in stack]).view(len(stack), -1)
                assert torch.equal(ref, torch.from_numpy(lv[k])), (image_size, k)
