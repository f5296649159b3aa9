"""ctypes binding of libmmgt_b200.so (the C ABI declared in include/mmgt_b200.h).

The product path has NO fallback: if the shared library is missing, or an entry point returns a
non-zero status, this module raises.  Nothing here imports ``oracle``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmgt_b200.so")

F32, BF16 = 0, 1

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class GemmParams(C.Structure):
    _fields_ = [("A", c_void_p), ("W", c_void_p), ("D", c_void_p), ("bias", c_void_p), ("rowscale", c_void_p),
                ("rowbias", c_void_p), ("residual", c_void_p),
                ("lda", c_int64), ("ldw", c_int64), ("ldd", c_int64), ("ldr", c_int64),
                ("M", c_int), ("N", c_int), ("K", c_int), ("rows_per_group", c_int), ("alpha", c_float),
                ("geglu_block", c_int), ("dtype", c_int), ("out_f32", c_int), ("exchange", c_void_p),
                ("ld_rowbias", c_int64), ("rowbias_mod", c_int), ("rowstats", c_void_p), ("colsum", c_void_p), ("act", c_int)]


class Conv3x3Params(C.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("y", c_void_p), ("bias", c_void_p), ("rowbias", c_void_p),
                ("residual", c_void_p),
                ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int), ("stride", c_int),
                ("upsample2x", c_int), ("frames_per_group", c_int), ("dtype", c_int), ("ld_rowbias", c_int64), ("act", c_int), ("w_subpixel", c_void_p)]


class AttentionParams(C.Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("k2", c_void_p), ("v2", c_void_p),
                ("seg2_index", c_void_p), ("out", c_void_p),
                ("ldq", c_int64), ("ldk", c_int64), ("ldv", c_int64), ("ldk2", c_int64), ("ldv2", c_int64),
                ("ldo", c_int64), ("kv_batch_stride", c_int64),
                ("N", c_int), ("Lq", c_int), ("Lk", c_int), ("Lk2", c_int), ("heads", c_int), ("d", c_int), ("B2", c_int),
                ("scale", c_float), ("dtype", c_int)]


class AudioAttentionParams(C.Structure):
    _fields_ = [("q3", c_void_p), ("kv6", c_void_p), ("mask", c_void_p * 3), ("scale", c_float * 3), ("out", c_void_p),
                ("ldq", c_int64), ("ldkv", c_int64), ("ldo", c_int64),
                ("N", c_int), ("T", c_int), ("M", c_int), ("heads", c_int), ("d", c_int),
                ("softmax_scale", c_float), ("dtype", c_int)]


MAX_PEERS = 8


class RowExchange(C.Structure):
    """mmgt_row_exchange: frame-shard <-> token-shard row mapping + the peers' receive buffers."""
    _fields_ = [("peer_base", c_void_p * MAX_PEERS), ("k", c_int), ("my", c_int), ("direction", c_int),
                ("B", c_int), ("F", c_int), ("T", c_int), ("ld", c_int64)]


class PeerBarrierParams(C.Structure):
    _fields_ = [("flags", c_void_p * MAX_PEERS), ("epoch", c_void_p), ("status", c_void_p), ("k", c_int), ("my", c_int),
                ("timeout_ms", c_int)]


# name -> (restype, argtypes); must list every symbol include/mmgt_b200.h declares (tests check this)
SIGNATURES = {
    "mmgt_abi_version": (c_int, []),
    "mmgt_ctx_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "mmgt_ctx_destroy": (c_int, [c_void_p]),
    "mmgt_last_error": (C.c_char_p, []),
    "mmgt_ctx_flag": (c_int64, [c_void_p, c_int, c_int64]),
    "mmgt_ncfhw_to_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "mmgt_tokens_to_ncfhw": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "mmgt_groupnorm": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_float, c_int, c_int, c_void_p]),
    "mmgt_layernorm": (c_int, [c_void_p] * 6 + [c_int64, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "mmgt_row_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int64, c_float, c_int, c_void_p]),
    "mmgt_gemm": (c_int, [c_void_p, C.POINTER(GemmParams), c_void_p]),
    "mmgt_gemm_tc_block_n": (c_int, [c_int]),
    "mmgt_conv3x3_workspace_bytes": (c_int64, [c_void_p, C.POINTER(Conv3x3Params)]),
    "mmgt_conv3x3": (c_int, [c_void_p, C.POINTER(Conv3x3Params), c_void_p, c_int64, c_void_p]),
    "mmgt_attention": (c_int, [c_void_p, C.POINTER(AttentionParams), c_void_p]),
    "mmgt_audio_attention": (c_int, [c_void_p, C.POINTER(AudioAttentionParams), c_void_p]),
    "mmgt_temporal_attention": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_float, c_int, c_void_p]),
    "mmgt_timestep_embedding": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "mmgt_silu_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mmgt_upsample_nearest2x": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "mmgt_im2col3x3": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "mmgt_pad_channels": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "mmgt_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "mmgt_window_accumulate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "mmgt_cfg_ddim_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 4 + [c_float] * 3 + [c_void_p]),
    "mmgt_mask_resize": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_float, c_void_p]),
    "mmgt_peer_alloc": (c_int, [c_void_p, c_int64, C.POINTER(c_void_p)]),
    "mmgt_peer_free": (c_int, [c_void_p, c_void_p]),
    "mmgt_peer_export": (c_int, [c_void_p, c_void_p, C.c_char_p]),
    "mmgt_peer_import": (c_int, [c_void_p, C.c_char_p, C.POINTER(c_void_p)]),
    "mmgt_peer_unmap": (c_int, [c_void_p, c_void_p]),
    "mmgt_peer_barrier": (c_int, [c_void_p, C.POINTER(PeerBarrierParams), c_void_p]),
    "mmgt_row_exchange_copy": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, C.POINTER(RowExchange), c_void_p]),
}

_lib = None


def load_library() -> C.CDLL:
    """Loads the shared library (no GPU needed); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m mmgt_b200.build` (there is no CPU / PyTorch "
                "fallback for the mmgt_b200 kernels)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class MmgtError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().mmgt_last_error().decode(errors="replace")
        raise MmgtError(f"{what} failed with status {rc}: {msg}")


class Context:
    """Per-device kernel context (opaque mmgt_ctx*)."""

    _by_device = {}

    def __init__(self, device_index: int):
        lib = load_library()
        h = c_void_p()
        check(lib.mmgt_ctx_create(C.byref(h), int(device_index)), "mmgt_ctx_create")
        self.handle = h
        self.device_index = device_index
        self.lib = lib
        if os.environ.get("MMGT_PDL") is not None:          # A/B switch for tests and profiling
            lib.mmgt_ctx_flag(h, 3, 1 if os.environ["MMGT_PDL"] not in ("0", "") else 0)

    @classmethod
    def get(cls, device_index: int) -> "Context":
        if device_index not in cls._by_device:
            cls._by_device[device_index] = cls(device_index)
        return cls._by_device[device_index]

    def set_tensor_cores(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 0, 1 if on else 0))

    def tensor_cores(self) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 0, -1))

    def set_resident_weights(self, on: bool) -> bool:
        """Weight-stationary GEMM variant (small K); on by default, switchable for A/B timing and tests."""
        return bool(self.lib.mmgt_ctx_flag(self.handle, 2, 1 if on else 0))

    def set_pdl(self, on: bool) -> bool:
        """Programmatic dependent launch (kernel i+1 is scheduled while kernel i drains); see common.cuh."""
        return bool(self.lib.mmgt_ctx_flag(self.handle, 3, 1 if on else 0))

    def pdl(self) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 3, -1))

    def launches(self) -> int:
        return int(self.lib.mmgt_ctx_flag(self.handle, 1, -1))

    def set_strict_tensor_cores(self, on: bool) -> bool:
        """bf16 requests without a tensor-core kernel raise (MMGT_E_UNSUPPORTED) instead of running on CUDA cores."""
        return bool(self.lib.mmgt_ctx_flag(self.handle, 4, 1 if on else 0))

    def simt_launches(self) -> int:
        """bf16 operator calls that took a CUDA-core kernel although tensor cores were enabled (shape cliffs)."""
        return int(self.lib.mmgt_ctx_flag(self.handle, 5, -1))

    def set_geglu_exact(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 6, 1 if on else 0))

    def set_groupnorm_split(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 7, 1 if on else 0))

    def set_residual_mma(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 13, 1 if on else 0))

    def set_tma_store(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 12, 1 if on else 0))

    def set_lean_epilogue(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 11, 1 if on else 0))

    def set_temporal_rows(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 10, 1 if on else 0))

    def set_attention_persistent(self, on) -> int:
        """False / 0: one item per CTA; True / 1: persistent CTAs when there are more items than SMs; n >= 2: always, n CTAs."""
        return int(self.lib.mmgt_ctx_flag(self.handle, 14, int(on)))

    def set_attention_q256(self, on) -> int:
        """Head dim <= 64: 256 queries per CTA, one query tile and one MMA-issuing warp per softmax group (flag 15).
        0 off; 1 = every exponential on MUFU; 5 = 1 of 4 score pairs through the FMA-pipe exp2; 4 = 2 of 4;
        8 / 7 / 9 = 1 / 5 / 4 with the softmax row sums from the tensor cores (P x ones into spare accumulator columns)."""
        return int(self.lib.mmgt_ctx_flag(self.handle, 15, int(on)))

    def set_layernorm_persistent(self, mode) -> int:
        """LayerNorm grid (flag 17): 0 = up to 16 blocks per SM, 1 = one exact wave of persistent blocks, 2 = + next-row prefetch."""
        return int(self.lib.mmgt_ctx_flag(self.handle, 17, int(mode)))

    def set_attention_packed(self, on: bool) -> bool:
        """Attention softmax loops on packed fp32 pairs (FFMA2 / FADD2), bit-identical to the scalar form (flag 16)."""
        return bool(self.lib.mmgt_ctx_flag(self.handle, 16, 1 if on else 0))

    def set_attention_v2(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 9, 1 if on else 0))

    def set_conv_implicit_all(self, on: bool) -> bool:
        return bool(self.lib.mmgt_ctx_flag(self.handle, 8, 1 if on else 0))
