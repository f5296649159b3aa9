"""Frame-sharded execution of one context window over k GPUs (level 3).

Every layer of the UNet except the 21 motion modules is per-frame, so k ranks each run F/k of a window's F
frames.  Around a motion module the rows switch from *frame-sharded* ``(b, f_loc, t)`` to *token-sharded*
``(b, f, t_loc)`` (all F frames of T/k pixels) and back -- the reference's ``(b f) d c <-> (b d) f c``
rearranges (motion_module.py:361-363,386) become an all-to-all.  Here the all-to-all has no kernel of its
own: the GEMM that produces the rows (``proj_in`` on the way in, the feed-forward's output projection on the
way back) stores each row straight into the receive buffer of the shard that owns it, over NVLink
(``mmgt_row_exchange`` in include/mmgt_b200.h), and a 32-thread flag barrier (``mmgt_peer_barrier``) orders
producers and consumers.

Buffers: every shard owns two receive buffers that alternate between exchanges.  Exchange n+2 reuses the
buffer of exchange n; a rank starts writing exchange n+2 only after it passed barrier n+1, which every peer
signalled after its own GEMM n+1 -- i.e. after it finished reading what exchange n delivered (the received
rows are last read by the first attention block's residual / by ``proj_out``, both before the next producing
GEMM in stream order).
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import PeerBarrierParams, RowExchange, check

FLAG_BYTES = 4096          # 8 flag slots, then epoch (word 64) and status (word 65)
_EPOCH_WORD, _STATUS_WORD = 64, 65


def exchange_destination(direction: int, m: int, k: int, my: int, B: int, F: int, T: int):
    """Host mirror of ``exchange_row_ptr`` (csrc/common.cuh): source row m -> (shard, destination row)."""
    Fl, Tc = F // k, T // k
    if direction == 1:
        n_loc, t = divmod(m, T)
        b, f_loc = divmod(n_loc, Fl)
        s = t // Tc
        return s, (b * F + my * Fl + f_loc) * Tc + (t - s * Tc)
    n, t_loc = divmod(m, Tc)
    b, f = divmod(n, F)
    s = f // Fl
    return s, (b * Fl + (f - s * Fl)) * T + my * Tc + t_loc


class _RawDeviceBytes:
    """CUDA-array-interface view of a raw device allocation, so torch can wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)


class Exchange:
    """One row exchange: the C descriptor + the local receive buffer viewed as (rows, C) in the run dtype."""

    def __init__(self, desc: RowExchange, recv: torch.Tensor):
        self.desc = desc
        self.recv = recv
        self.recv_dummy = recv


class FrameShardGroup:
    """The k ranks that share one window.  ``shard`` = this rank's index inside the group."""

    def __init__(self, eng, k: int, shard: int, recv_ptrs: List[List[int]], flag_ptrs: List[int], buf_bytes: int,
                 owned: Optional[list] = None, mapped: Optional[list] = None):
        if not (1 <= k <= _lib.MAX_PEERS):
            raise ValueError(f"frame shards must be 1..{_lib.MAX_PEERS}, got {k}")
        self.eng, self.k, self.shard, self.buf_bytes = eng, k, shard, buf_bytes
        self.recv_ptrs, self.flag_ptrs = recv_ptrs, flag_ptrs            # [buffer][shard], [shard]
        self._owned, self._mapped = owned or [], mapped or []
        self._local = [torch.as_tensor(_RawDeviceBytes(recv_ptrs[i][shard], buf_bytes), device=eng.device) for i in range(2)]
        self._n = 0
        self._cache = {}
        bp = PeerBarrierParams()
        for s in range(k):
            bp.flags[s] = flag_ptrs[s]
        bp.epoch = flag_ptrs[shard] + 4 * _EPOCH_WORD
        bp.status = flag_ptrs[shard] + 4 * _STATUS_WORD
        bp.k, bp.my, bp.timeout_ms = k, shard, 30000
        self._barrier = bp
        self._status = torch.as_tensor(_RawDeviceBytes(flag_ptrs[shard] + 4 * _STATUS_WORD, 4), device=eng.device).view(torch.int32)

    # ------------------------------------------------------------------ construction
    @classmethod
    def create(cls, eng, rank: int, world: int, k: int, buf_bytes: int, process_group=None) -> "FrameShardGroup":
        """Collective over ``process_group`` (all ``world`` ranks): allocate, export, all-gather and import the
        IPC handles.  Rank r is shard r % k of group r // k."""
        import torch.distributed as dist
        if world % k:
            raise ValueError(f"world size {world} is not a multiple of frame_shards={k}")
        lib, h = eng.lib, eng.h
        owned = []

        def alloc(nbytes):
            p = C.c_void_p()
            check(lib.mmgt_peer_alloc(h, nbytes, C.byref(p)), "mmgt_peer_alloc")
            owned.append(p.value)
            return p.value
        mine = [alloc(buf_bytes), alloc(buf_bytes), alloc(FLAG_BYTES)]
        blob = b""
        for p in mine:
            hb = C.create_string_buffer(64)
            check(lib.mmgt_peer_export(h, C.c_void_p(p), hb), "mmgt_peer_export")
            blob += hb.raw
        gathered = [None] * world
        dist.all_gather_object(gathered, blob, group=process_group)
        g0 = (rank // k) * k
        shard = rank - g0
        recv_ptrs, flag_ptrs, mapped = [[0] * k, [0] * k], [0] * k, []
        error = None
        try:
            for s in range(k):
                if s == shard:
                    ptrs = mine
                else:
                    ptrs = []
                    for i in range(3):
                        p = C.c_void_p()
                        check(lib.mmgt_peer_import(h, gathered[g0 + s][64 * i:64 * (i + 1)], C.byref(p)),
                              f"mmgt_peer_import (shard {s}: CUDA IPC / NVLink peer access is required for frame sharding)")
                        mapped.append(p.value)
                        ptrs.append(p.value)
                recv_ptrs[0][s], recv_ptrs[1][s], flag_ptrs[s] = ptrs
        except Exception as e:   # noqa: BLE001 -- every rank must learn about it: the others are about to wait for us
            error = e
        ok = torch.tensor([0.0 if error is not None else 1.0], device=eng.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
        if float(ok) != 1.0:
            for p in mapped:
                lib.mmgt_peer_unmap(h, C.c_void_p(p))
            dist.barrier(group=process_group)
            for p in owned:
                lib.mmgt_peer_free(h, C.c_void_p(p))
            raise RuntimeError(f"frame-shard peer memory could not be mapped on every rank ({error or 'a peer failed'})")
        return cls(eng, k, shard, recv_ptrs, flag_ptrs, buf_bytes, owned, mapped)

    @classmethod
    def emulate(cls, eng, k: int, buf_bytes: int) -> List["FrameShardGroup"]:
        """k shards inside ONE process on one device (tests of the row mapping and of the flag protocol)."""
        lib, h = eng.lib, eng.h
        owned = []

        def alloc(nbytes):
            p = C.c_void_p()
            check(lib.mmgt_peer_alloc(h, nbytes, C.byref(p)), "mmgt_peer_alloc")
            owned.append(p.value)
            return p.value
        recv_ptrs = [[alloc(buf_bytes) for _ in range(k)] for _ in range(2)]
        flag_ptrs = [alloc(FLAG_BYTES) for _ in range(k)]
        groups = [cls(eng, k, s, recv_ptrs, flag_ptrs, buf_bytes) for s in range(k)]
        groups[0]._owned = owned
        return groups

    def close(self, process_group=None):
        """Unmap the peers' buffers, then (after a barrier when the group spans processes) free the own ones."""
        lib, h = self.eng.lib, self.eng.h
        torch.cuda.synchronize()
        self._local, self._cache, self._status = [], {}, None
        for p in self._mapped:
            lib.mmgt_peer_unmap(h, C.c_void_p(p))
        if self._mapped:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.barrier(group=process_group)
        for p in self._owned:
            lib.mmgt_peer_free(h, C.c_void_p(p))
        self._mapped, self._owned = [], []

    # ------------------------------------------------------------------ per exchange
    def exchange(self, direction: int, B: int, F: int, T: int, Cc: int) -> Exchange:
        """Descriptor of the next exchange (alternates the receive buffer).  F = frames of the whole window."""
        k = self.k
        if F % k or T % k:
            raise ValueError(f"frame sharding needs frames ({F}) and tokens per frame ({T}) divisible by {k}")
        buf = self._n & 1
        self._n += 1
        key = (direction, B, F, T, Cc, buf)
        ex = self._cache.get(key)
        if ex is None:
            esize = torch.empty((), dtype=self.eng.dtype).element_size()
            rows = B * (F // k) * T
            if rows * Cc * esize > self.buf_bytes:
                raise ValueError(f"receive buffer too small: {rows}x{Cc} rows need {rows * Cc * esize} B > {self.buf_bytes} B")
            d = RowExchange()
            for s in range(k):
                d.peer_base[s] = self.recv_ptrs[buf][s]
            d.k, d.my, d.direction, d.B, d.F, d.T, d.ld = k, self.shard, direction, B, F, T, Cc
            recv = self._local[buf][: rows * Cc * esize].view(self.eng.dtype).view(rows, Cc)
            ex = Exchange(d, recv)
            self._cache[key] = ex
        return ex

    def barrier(self):
        from .kernels import _stream
        check(self.eng.lib.mmgt_peer_barrier(self.eng.h, C.byref(self._barrier), _stream()), "mmgt_peer_barrier")

    def check(self):
        """Synchronises; raises if any barrier of this rank timed out (a peer died or ran a different schedule)."""
        torch.cuda.synchronize()
        if int(self._status.item()) != 0:
            raise RuntimeError(f"frame-shard barrier timed out on shard {self.shard} of {self.k}")


def max_exchange_bytes(B: int, frames_local: int, latent_tokens: int, base_channels: int, esize: int) -> int:
    """Largest exchanged tensor = the finest level: B * F/k frames * T tokens * C channels."""
    return B * frames_local * latent_tokens * base_channels * esize
