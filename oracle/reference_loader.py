"""Import the reference's hot-path modules UNCHANGED (only possible where /root/reference exists).

TEST INFRASTRUCTURE.  Used by ``make_golden.py`` and by ``tests/test_oracle_vs_reference.py``
(skipped on machines without the reference checkout, e.g. the GPU box).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("MMGT_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")

# stable-diffusion-v1-5/unet/config.json values the reference reads
SD15_CFG = dict(
    sample_size=64, in_channels=4, out_channels=4, center_input_sample=False, flip_sin_to_cos=True,
    freq_shift=0, block_out_channels=[320, 640, 1280, 1280], layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768,
    attention_head_dim=8,
)
# config/prompts/animation.yaml:47-75
UNET_ADDITIONAL_KWARGS = dict(
    use_inflated_groupnorm=True, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False,
    use_motion_module=True, use_audio_module=True, motion_module_resolutions=[1, 2, 4, 8],
    motion_module_mid_block=True, motion_module_decoder_only=False, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=["Temporal_Self", "Temporal_Self"],
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1),
    audio_attention_dim=768, stack_enable_blocks_name=["up", "down", "mid"],
    stack_enable_blocks_depth=[0, 1, 2, 3],
)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _activate():
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, SHIM_ROOT):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, SHIM_ROOT)


def load_reference_modules():
    """Returns the reference's module namespaces (unet_3d, mutual_self_attention, context)."""
    _activate()
    import importlib

    unet_3d = importlib.import_module("src.models.unet_3d")
    msa = importlib.import_module("src.models.mutual_self_attention")
    context = importlib.import_module("src.pipelines.context")
    attention = importlib.import_module("src.models.attention")
    return dict(unet_3d=unet_3d, mutual_self_attention=msa, context=context, attention=attention)


def build_reference_unet(block_out_channels=None, cfg_overrides=None):
    mods = load_reference_modules()
    cfg = dict(SD15_CFG)
    if block_out_channels is not None:
        cfg["block_out_channels"] = list(block_out_channels)
    if cfg_overrides:
        cfg.update(cfg_overrides)
    unet = mods["unet_3d"].UNet3DConditionModel.from_config(cfg, **UNET_ADDITIONAL_KWARGS)
    return unet, mods
