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


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)], ids=["f32", "bf16tc"])
def test_reference_net_write_pass_matches_reference_banks(dtype, tol):  # bf16: bank taps sit behind up to ~60 bf16-stored layers, same bound as the forward output (TOL_BF16_FWD)
    """ReferenceNet write pass on the kernels: 16 banks vs the reference's own write-mode controller + blocks."""
    from test_host_mirror_cpu import _check_refnet_banks, _tiny_refnet
    net = _tiny_refnet("cuda")
    net.set_compute_dtype(dtype)
    e_bank, e_out = _check_refnet_banks(net, tol)
    print(f"ReferenceNet write pass {dtype}: worst bank rel-L2 {e_bank:.3e}, output {e_out:.3e}")


def test_pipeline_with_kernel_reference_net_and_pose_guider_matches_oracle_loop():
    """Pose2VideoPipeline.__call__ with PIL inputs where the ReferenceNet AND the PoseGuider are this package's classes (f1) and
    the denoising UNet runs in float32: CLIP / VAE are stubs; the banks the reader receives must equal the write-pass banks
    and the loop must reproduce the oracle fed with them."""
    from PIL import Image
    from helpers import TINY, bank_pairing_order, build_cuda_unet, synthetic_state_dict
    from mmgt_b200.mutual_self_attention import ReferenceAttentionControl, _writer_blocks
    from mmgt_b200.pipeline_pose2vid_long import Pose2VideoPipeline
    from mmgt_b200.pose_guider import PoseGuider
    from mmgt_b200.scheduling_ddim import DDIMSchedule
    from oracle.sampler import DDIM, denoise_step, uniform_windows
    from oracle.synthetic import make_inputs
    from oracle.unet3d import UNetSpec, unet3d_forward
    from test_host_mirror_cpu import _StubClip, _StubVae, _tiny_refnet
    torch.manual_seed(3)
    spec = UNetSpec(block_out_channels=TINY)
    sd = synthetic_state_dict("tiny")
    unet = build_cuda_unet(TINY, sd, compute_dtype=torch.float32)
    unet.train()
    unet.enable_gradient_checkpointing()
    refnet = _tiny_refnet("cuda").set_compute_dtype(torch.float32)
    guider = PoseGuider(conditioning_embedding_channels=TINY[0]).cuda()
    with torch.no_grad():
        for p in guider.conv_out.parameters():
            p.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(4)).cuda() * 0.02)
    vae, clip = _StubVae().cuda(), _StubClip().cuda()
    pipe = Pose2VideoPipeline(vae=vae, image_encoder=clip, reference_unet=refnet, denoising_unet=unet, pose_guider=guider,
                              scheduler=DDIMSchedule.from_config())
    L, latent, n_steps = 16, 16, 2
    size = latent * 8
    rng = np.random.default_rng(1)
    pil = lambda: Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8), "RGB")   # noqa: E731
    ref_image, poses = pil(), [pil() for _ in range(L)]
    inp = make_inputs(spec, L, latent, seed=3)
    cond = lambda ms: [m[L:] for m in ms]   # noqa: E731
    audio = inp["audio"][1:2]
    out = pipe(ref_image=ref_image, pose_images=poses, audio_tensor=audio, pixel_values_full_mask=cond(inp["full_mask"]),
               pixel_values_face_mask=cond(inp["face_mask"]), pixel_values_lip_mask=cond(inp["lip_mask"]), width=size,
               height=size, video_length=L, num_inference_steps=n_steps, guidance_scale=3.5,
               generator=torch.Generator().manual_seed(42), motion_scale=[1.0, 1.0, 2.0], output_type="latent").videos
    # the same conditioning by hand -> oracle loop
    with torch.no_grad():
        from transformers import CLIPImageProcessor
        px = CLIPImageProcessor().preprocess(ref_image.resize((224, 224)), return_tensors="pt").pixel_values.cuda()
        e = clip(px).image_embeds.unsqueeze(1)
        ehs = torch.cat([torch.zeros_like(e), e])
        ref_lat = vae.encode(pipe.ref_image_processor.preprocess(ref_image, height=size, width=size).cuda()).latent_dist.mean * 0.18215
        writer = ReferenceAttentionControl(refnet, do_classifier_free_guidance=True, mode="write", fusion_blocks="full")
        refnet(ref_lat.repeat(2, 1, 1, 1), torch.zeros((), device="cuda", dtype=torch.long), encoder_hidden_states=ehs,
               return_dict=False)
        feats = [b.bank[0].half().float().cpu() for b in _writer_blocks(refnet, "full")]
        writer.clear()
        writer.remove()
        banks = dict(zip(bank_pairing_order(spec), feats))
        pose_fea = guider(torch.cat([pipe.cond_image_processor.preprocess(p, height=size, width=size).unsqueeze(2)
                                     for p in poses], dim=2).cuda()).float().cpu()
    dup = lambda ms: [torch.cat([m, m]) for m in cond(ms)]   # noqa: E731

    def unet_fn(sample, t, ehs_, aud, pose, full, face, lip, ms):
        with torch.no_grad():
            return unet3d_forward(sd, spec, sample, t, ehs_, aud, pose, full, face, lip, ms, banks, ref_index=[None, 1],
                                  apply_motion_scale=True)
    ddim = DDIM()
    lat = torch.randn((1, 4, L, latent, latent), generator=torch.Generator().manual_seed(42))
    for t in ddim.timesteps(n_steps):
        lat, _ = denoise_step(unet_fn, lat, t, n_steps, ddim, 3.5, uniform_windows(0, L), pose_fea,
                              torch.cat([torch.zeros_like(audio), audio]), dup(inp["full_mask"]), dup(inp["face_mask"]),
                              dup(inp["lip_mask"]), ehs.float().cpu(), [1.0, 1.0, 2.0])
    err = rel_l2(out, lat)
    print(f"pipeline with kernel ReferenceNet + PoseGuider vs oracle loop: latents rel-L2 {err:.3e}")
    assert err < 1e-4
