"""Host mirror of src/models/mutual_self_attention.py (read side).

The reference monkey-patches ``forward`` of every spatial transformer block and hangs a ``bank`` list on it.
Here the blocks implement the read-mode computation natively, so the controller only (a) records the CFG
layout on the UNet, (b) pairs reader / writer blocks in the reference's order and copies the banks
(fp16-rounded by default, mutual_self_attention.py:304,340), (c) clears them.
The writer may be the reference's own 2-D ReferenceNet under the reference's controller in "write" mode:
pairing is by duck typing (modules that own ``bank`` and ``norm1``).
"""
import torch

from .attention import TemporalBasicTransformerBlock


def torch_dfs(model: torch.nn.Module):
    result = [model]
    for child in model.children():
        result += torch_dfs(child)
    return result


def _reader_blocks(unet, fusion_blocks):
    if fusion_blocks == "midup":
        mods = torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks)
    else:
        mods = torch_dfs(unet)
    blocks = [m for m in mods if isinstance(m, TemporalBasicTransformerBlock)]
    return sorted(blocks, key=lambda m: -m.norm1.normalized_shape[0])      # stable, like the reference


def _writer_blocks(unet, fusion_blocks):
    if fusion_blocks == "midup":
        mods = torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks)
    else:
        mods = torch_dfs(unet)
    blocks = [m for m in mods if hasattr(m, "bank") and hasattr(m, "norm1")
              and type(m).__name__ in ("BasicTransformerBlock", "TemporalBasicTransformerBlock")]
    return sorted(blocks, key=lambda m: -m.norm1.normalized_shape[0])


class ReferenceAttentionControl:
    def __init__(self, unet, mode="write", do_classifier_free_guidance=False, attention_auto_machine_weight=float("inf"),
                 gn_auto_machine_weight=1.0, style_fidelity=1.0, reference_attn=True, reference_adain=False,
                 fusion_blocks="midup", batch_size=1) -> None:
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        if mode != "read":
            raise NotImplementedError("mmgt_b200 implements the read side (denoising UNet); the ReferenceNet write "
                                      "pass stays the reference's PyTorch (SURVEY section 8 f1)")
        if reference_adain:
            raise NotImplementedError("reference_adain")
        self.unet = unet
        self.reference_attn = reference_attn
        self.fusion_blocks = fusion_blocks
        self.do_classifier_free_guidance = do_classifier_free_guidance
        if reference_attn:
            blocks = _reader_blocks(unet, fusion_blocks)
            for i, m in enumerate(blocks):
                m.bank = []
                m.attn_weight = float(i) / float(len(blocks))
            unet._reference_control = dict(do_classifier_free_guidance=do_classifier_free_guidance,
                                           fusion_blocks=fusion_blocks)

    def update(self, writer, dtype=torch.float16):
        if not self.reference_attn:
            return
        readers = _reader_blocks(self.unet, self.fusion_blocks)
        writers = _writer_blocks(writer.unet, self.fusion_blocks)
        for r, w in zip(readers, writers):
            r.bank = [v.clone().to(dtype) for v in w.bank]

    def set_banks(self, banks, dtype=torch.float16):
        """Direct form of ``update``: ``banks`` = list of (Bb, T, C) tensors in pairing order."""
        readers = _reader_blocks(self.unet, self.fusion_blocks)
        assert len(banks) == len(readers)
        for r, b in zip(readers, banks):
            r.bank = [b.clone().to(dtype)]

    def clear(self):
        if self.reference_attn:
            for r in _reader_blocks(self.unet, self.fusion_blocks):
                r.bank.clear()
                r._bank_kv = None
