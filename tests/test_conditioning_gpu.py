"""GPU: the one-shot conditioning modules on the kernel set against outputs of the reference's own
PoseGuider / AudioProjModel (tests/golden/conditioning.npz, oracle/make_golden_f1.py), and against the oracle restatement
at the real 512x512 shape."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import GOLD, rel_l2  # noqa: E402
from oracle.conditioning import pose_guider_forward  # noqa: E402
from oracle.make_golden_f1 import POSE_CFG, audio_input, pose_input  # noqa: E402
from test_host_mirror_cpu import _conditioning_modules  # noqa: E402


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)], ids=["f32", "bf16tc"])
def test_pose_guider_and_audio_proj_match_reference_golden(dtype, tol):
    g = np.load(os.path.join(GOLD, "conditioning.npz"))
    pg, ap = _conditioning_modules()
    pg.to("cuda")
    ap.to("cuda")
    pg.compute_dtype = ap.compute_dtype = dtype
    from mmgt_b200.kernels import get_engine
    eng = get_engine(torch.device("cuda", 0), dtype)
    n_simt = eng.ctx.simt_launches()
    y = pg(pose_input().cuda())
    z = ap(audio_input().cuda())
    e_p, e_a = rel_l2(y, torch.from_numpy(g["pose_out"])), rel_l2(z, torch.from_numpy(g["audio_out"]))
    print(f"PoseGuider {dtype}: {e_p:.3e}; AudioProjModel: {e_a:.3e}")
    assert y.shape == g["pose_out"].shape and z.shape == g["audio_out"].shape
    assert e_p < tol and e_a < tol
    if dtype == torch.bfloat16:      # channel padding in the packs keeps all eight convolutions on the tensor cores
        assert eng.ctx.simt_launches() == n_simt


def test_pose_guider_512_bf16_vs_oracle():
    """The real shape: 4 pose frames at 512x512 -> (1, 320, 4, 64, 64)."""
    pg, _ = _conditioning_modules()
    x = torch.rand(1, 3, 4, 512, 512, generator=torch.Generator().manual_seed(1))
    sd = {k: v.detach().clone() for k, v in pg.state_dict().items()}
    with torch.no_grad():
        ref = pose_guider_forward(sd, x, POSE_CFG["block_out_channels"])
    pg.to("cuda")
    pg.compute_dtype = torch.bfloat16
    y = pg(x.cuda())
    assert y.shape == (1, 320, 4, 64, 64)
    assert rel_l2(y, ref) < 1e-2
