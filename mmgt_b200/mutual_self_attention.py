"""Host mirror of src/models/mutual_self_attention.py.

The reference monkey-patches ``forward`` of every spatial transformer block and hangs a ``bank`` list on it.
Read side (the denoising UNet of this package): the blocks implement the read-mode computation natively, so the
controller only (a) records the CFG layout on the UNet, (b) pairs reader / writer blocks in the reference's order and
copies the banks (fp16-rounded by default, mutual_self_attention.py:304,340), (c) clears them.
Write side (the ReferenceNet, any 2-D UNet whose transformer blocks own ``norm1``: the reference's PyTorch
``UNet2DConditionModel``): what write mode stores is ``norm1(hidden_states)`` at the entry of every block
(mutual_self_attention.py:139-148), so instead of replacing ``forward`` the controller registers a forward pre-hook
that appends exactly that tensor to ``module.bank``; the block's own forward then runs unchanged (the hacked write
forward computes the same thing as the original one).  A writer controlled by the reference's own
``ReferenceAttentionControl`` works too: pairing is by duck typing (modules that own ``bank`` and ``norm1``).
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


def _writer_blocks(unet, fusion_blocks, need_bank=True):
    if fusion_blocks == "midup":
        mods = torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks)
    else:
        mods = torch_dfs(unet)
    blocks = [m for m in mods if (hasattr(m, "bank") or not need_bank) and hasattr(m, "norm1")
              and type(m).__name__ in ("BasicTransformerBlock", "TemporalBasicTransformerBlock")]
    return sorted(blocks, key=lambda m: -m.norm1.normalized_shape[0])


def _bank_write_hook(module, args, kwargs=None):
    """Write mode (mutual_self_attention.py:139-141): bank.append(norm1(hidden_states).clone())."""
    hidden_states = args[0] if args else (kwargs or {}).get("hidden_states")
    with torch.no_grad():
        module.bank.append(module.norm1(hidden_states).clone())
    return None


class ReferenceAttentionControl:
    def __init__(self, unet, mode="write", do_classifier_free_guidance=False, attention_auto_machine_weight=float("inf"),
                 gn_auto_machine_weight=1.0, style_fidelity=1.0, reference_attn=True, reference_adain=False,
                 fusion_blocks="midup", batch_size=1) -> None:
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        if reference_adain:
            raise NotImplementedError("reference_adain")
        self.unet = unet
        self.mode = mode
        self.reference_attn = reference_attn
        self.fusion_blocks = fusion_blocks
        self.do_classifier_free_guidance = do_classifier_free_guidance
        self._hooks = []
        if reference_attn and mode == "write":
            blocks = _writer_blocks(unet, fusion_blocks, need_bank=False)
            if not blocks:
                raise ValueError("write mode: the UNet has no BasicTransformerBlock / TemporalBasicTransformerBlock with norm1")
            for i, m in enumerate(blocks):
                m.bank = []
                m.attn_weight = float(i) / float(len(blocks))
                if isinstance(m, TemporalBasicTransformerBlock):
                    m.write_bank = True          # this package's ReferenceNet (unet_2d_condition.py): banked inside run()
                else:
                    self._hooks.append(m.register_forward_pre_hook(_bank_write_hook))
        elif reference_attn:
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
        if len(writers) != len(readers):
            raise ValueError(f"reference writer has {len(writers)} transformer blocks with a bank, the denoising UNet "
                             f"{len(readers)}: the two UNets must share the SD-1.5 block layout")
        for r, w in zip(readers, writers):
            if not w.bank:
                raise ValueError("reference writer bank is empty: run the ReferenceNet forward before update()")
            if w.bank[0].shape[-1] != r.norm1.normalized_shape[0]:
                raise ValueError(f"bank width {w.bank[0].shape[-1]} does not match the reader block width "
                                 f"{r.norm1.normalized_shape[0]}")
            r.bank = [v.clone().to(dtype) for v in w.bank]

    def set_banks(self, banks, dtype=torch.float16):
        """Direct form of ``update``: ``banks`` = list of (Bb, T, C) tensors in pairing order."""
        readers = _reader_blocks(self.unet, self.fusion_blocks)
        assert len(banks) == len(readers)
        for r, b in zip(readers, banks):
            r.bank = [b.clone().to(dtype)]

    def clear(self):
        if not self.reference_attn:
            return
        if self.mode == "write":
            for w in _writer_blocks(self.unet, self.fusion_blocks):
                w.bank.clear()
            return
        for r in _reader_blocks(self.unet, self.fusion_blocks):
            r.bank.clear()        # the projected K/V buffer stays allocated: a captured CUDA graph may still point at it

    def remove(self):
        """Write mode: take the pre-hooks / write flags off the ReferenceNet again."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self.mode == "write":
            for m in _writer_blocks(self.unet, self.fusion_blocks, need_bank=False):
                if isinstance(m, TemporalBasicTransformerBlock):
                    m.write_bank = False
