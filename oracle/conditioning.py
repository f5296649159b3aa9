"""CPU restatement of two one-shot conditioning modules of Pose2VideoPipeline.__call__.
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.

  pose_guider_forward  -- src/models/pose_guider.py:11-57 (PoseGuider: 8 InflatedConv3d + SiLU, 512^2 -> 64^2 x 320)
  audio_proj_forward   -- src/models/audio_proj.py:40-124 (AudioProjModel: 3 Linear + ReLU, LayerNorm -> 32 tokens x 768)

Both are functional (state dict in, tensors out) and pinned to outputs of the reference's own classes
(tests/golden/conditioning.npz, written by oracle/make_golden_f1.py; tests/test_oracle_golden.py).
"""
from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def _inflated_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """InflatedConv3d.forward (src/models/resnet.py:9-17): fold frames into the batch, Conv2d k=3 p=1, unfold."""
    B, C, Fr, H, W = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), w, b, stride=stride, padding=1)
    return y.reshape(B, Fr, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def pose_guider_forward(sd: Dict[str, torch.Tensor], conditioning: torch.Tensor,
                        block_out_channels: Sequence[int] = (16, 32, 64, 128)) -> torch.Tensor:
    """conditioning (B, 3, F, H, W) -> (B, C_emb, F, H/8, W/8).  pose_guider.py:47-57: conv_in, SiLU, then per level a
    stride-1 conv + SiLU and a stride-2 conv + SiLU (:26-36), then the zero-initialised conv_out (:38-45) without SiLU."""
    e = F.silu(_inflated_conv(conditioning, sd["conv_in.weight"], sd["conv_in.bias"]))
    for i in range(2 * (len(block_out_channels) - 1)):
        e = F.silu(_inflated_conv(e, sd[f"blocks.{i}.weight"], sd[f"blocks.{i}.bias"], stride=1 + (i % 2)))
    return _inflated_conv(e, sd["conv_out.weight"], sd["conv_out.bias"])


def audio_proj_forward(sd: Dict[str, torch.Tensor], audio_embeds: torch.Tensor, context_tokens: int = 32,
                       output_dim: int = 768) -> torch.Tensor:
    """audio_embeds (B, F, window, blocks, channels) -> (B, F, context_tokens, output_dim).  audio_proj.py:99-124:
    flatten (window, blocks, channels) per frame, proj1 + ReLU, proj2 + ReLU, proj3, reshape to tokens, LayerNorm."""
    B, Fr = audio_embeds.shape[:2]
    x = audio_embeds.reshape(B * Fr, -1)
    x = torch.relu(F.linear(x, sd["proj1.weight"], sd["proj1.bias"]))
    x = torch.relu(F.linear(x, sd["proj2.weight"], sd["proj2.bias"]))
    t = F.linear(x, sd["proj3.weight"], sd["proj3.bias"]).reshape(B * Fr, context_tokens, output_dim)
    t = F.layer_norm(t, (output_dim,), sd["norm.weight"], sd["norm.bias"], eps=1e-5)
    return t.reshape(B, Fr, context_tokens, output_dim)
