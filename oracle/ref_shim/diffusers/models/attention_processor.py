"""Attention + processors, restated from the published diffusers 0.24.0 behaviour (App. A)."""
from typing import Union

import torch
import torch.nn.functional as F
from torch import nn


class AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        q = attn.to_q(hidden_states)
        src = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(src), attn.to_v(src)
        b, lq, _ = q.shape
        h = attn.heads

        def split(t):
            return t.reshape(b, -1, h, t.shape[-1] // h).permute(0, 2, 1, 3).reshape(b * h, -1, t.shape[-1] // h)

        q, k, v = split(q), split(k), split(v)
        if attn.upcast_attention:
            q, k = q.float(), k.float()
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=attn.scale)
        probs = scores.softmax(dim=-1).to(v.dtype)
        out = torch.bmm(probs, v)
        out = out.reshape(b, h, lq, -1).permute(0, 2, 1, 3).reshape(b, lq, -1)
        out = attn.to_out[0](out)
        out = attn.to_out[1](out)
        return out / attn.rescale_output_factor


class AttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        src = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(src), attn.to_v(src)
        h = attn.heads
        d = k.shape[-1] // h
        q = q.view(b, -1, h, d).transpose(1, 2)
        k = k.view(b, -1, h, d).transpose(1, 2)
        v = v.view(b, -1, h, d).transpose(1, 2)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        out = out.transpose(1, 2).reshape(b, -1, h * d).to(q.dtype)
        out = attn.to_out[0](out)
        out = attn.to_out[1](out)
        return out / attn.rescale_output_factor


AttentionProcessor = Union[AttnProcessor, AttnProcessor2_0]


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, out_bias=True, scale_qk=True,
                 rescale_output_factor=1.0, processor=None, **unused):
        super().__init__()
        inner = dim_head * heads
        self.inner_dim = inner
        self.heads = heads
        self.upcast_attention = upcast_attention
        self.rescale_output_factor = rescale_output_factor
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor if processor is not None else AttnProcessor2_0()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
