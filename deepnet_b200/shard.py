"""Leading-axis sharding of large tensors over the GPUs of one box (new — the reference is single
device: `CudaContext(createNew=false)`, Tensor/Tensor/Cuda/CudaUtils.fs:42-46).

Host-side mirror of the `dn_shard_*` entry points of libdeepnet_b200.so (include/dn_tensor.h "Multi-GPU"): every rank
owns a contiguous slab of dim 0 of every operand; element-wise operators are the ordinary backend calls on each
device; reductions, arg-reductions, Find, TrueIndices and MaskedGet deliver the FULL result on every rank through
peer-memory stores over NVLink and a flag barrier inside the producing kernel — one launch per rank, no NCCL launch,
no torch.distributed. Two process models:

  * `ShardGroup.single_process(devices)` — one process drives every rank (the F# host's model);
  * `ShardGroup.multi_process(rank, world, device, exchange)` — one process per GPU (torchrun); `exchange(bytes) ->
    [bytes] * world` is any all-gather of 64-byte blobs the harness has (torch.distributed, MPI, files ...). It is
    used once, for the window handles.

Sharded results are allocated from the group's symmetric heap (`alloc`); all calls are collective. A thread that
drives several ranks issues each collective on all of them inside `with group.bracket():` (dn_shard_group_start /
dn_shard_group_end — the ncclGroupStart / ncclGroupEnd of this library).
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import dtypes, native
from . import layout as TL
from .backend import REDUCE_OPS, TensorCudaDevice
from .tensor import Tensor


def slab(nrows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [begin, begin+count) of `nrows` rows for `rank`; the remainder goes to the first ranks
    (dn_shard_slab)."""
    base, rem = divmod(nrows, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


_OUT_DTYPE = {"CountTrueLastAxis": dtypes.DN_I64, "ArgMinLastAxis": dtypes.DN_I64, "ArgMaxLastAxis": dtypes.DN_I64,
              "FindLastAxis": dtypes.DN_I64}
_ARG = {"ArgMinLastAxis": 0, "ArgMaxLastAxis": 1}


class ShardGroup:
    def __init__(self, dev: TensorCudaDevice, world: int, local: Dict[int, int], heap_bytes: int):
        self.dev = dev
        self.api = dev.api
        self.world = world
        self.local = dict(local)            # rank -> CUDA device
        self.ranks = sorted(self.local)
        g = C.c_void_p()
        n = len(self.ranks)
        rk = (C.c_int32 * n)(*self.ranks)
        dv = (C.c_int32 * n)(*[self.local[r] for r in self.ranks])
        self.api.call("shard_group_create", world, n, rk, dv, int(heap_bytes), C.byref(g))
        self._g = g

    # -- construction -------------------------------------------------------------------------------------------
    @staticmethod
    def single_process(dev: TensorCudaDevice, devices: Sequence[int], heap_bytes: int = 256 << 20) -> "ShardGroup":
        """One process drives rank r on CUDA device devices[r] (ranks may share a device)."""
        grp = ShardGroup(dev, len(devices), {r: d for r, d in enumerate(devices)}, heap_bytes)
        grp.api.call("shard_group_connect", grp._g, None)
        return grp

    @staticmethod
    def multi_process(dev: TensorCudaDevice, rank: int, world: int, device: int,
                      exchange: Callable[[bytes], List[bytes]], heap_bytes: int = 256 << 20) -> "ShardGroup":
        grp = ShardGroup(dev, world, {rank: device}, heap_bytes)
        h = C.create_string_buffer(native.DN_SHARD_HANDLE_BYTES)
        grp.api.call("shard_group_handle", grp._g, rank, h)
        blobs = exchange(h.raw)
        if len(blobs) != world or any(len(b) != native.DN_SHARD_HANDLE_BYTES for b in blobs):
            raise ValueError("exchange() must return one 64-byte handle per rank")
        grp.api.call("shard_group_connect", grp._g, C.create_string_buffer(b"".join(blobs), world * 64))
        return grp

    def close(self) -> None:
        if self._g:
            self.api.call("shard_group_destroy", self._g)
            self._g = None

    # -- plumbing -----------------------------------------------------------------------------------------------
    def set_stream(self, rank: int, stream: int) -> None:
        self.api.call("shard_set_stream", self._g, rank, stream)

    def sync(self, rank: Optional[int] = None) -> None:
        for r in (self.ranks if rank is None else [rank]):
            self.api.call("shard_sync", self._g, r)

    @contextmanager
    def bracket(self):
        """dn_shard_group_start / dn_shard_group_end around ONE collective issued on every local rank by this thread.
        A no-op when this process drives a single rank (every call then completes its own wait)."""
        if len(self.ranks) > 1:
            self.api.call("shard_group_start", self._g)
            try:
                yield
            finally:
                self.api.call("shard_group_end", self._g)
        else:
            yield

    @contextmanager
    def batch(self):
        """A bracket in any process model: the collectives issued inside launch back to back and share ONE wait at
        the end (dn_shard_group_end). Targets inside must be distinct."""
        self.api.call("shard_group_start", self._g)
        try:
            yield
        finally:
            self.api.call("shard_group_end", self._g)

    def barrier(self) -> None:
        with self.bracket():
            for r in self.ranks:
                self.api.call("shard_barrier", self._g, r)

    def heap_reset(self) -> None:
        for r in self.ranks:
            self.api.call("shard_heap_reset", self._g, r)

    def alloc(self, rank: int, shape: Sequence[int], dtype: int) -> Tensor:
        """A C-contiguous tensor in rank's symmetric heap (same offset on every rank when every rank allocates in
        the same order)."""
        shape = tuple(int(s) for s in shape)
        lay = TL.newC(shape)
        p = C.c_void_p()
        self.api.call("shard_heap_alloc", self._g, rank, max(1, lay.NElems) * dtypes.itemsize(dtype), C.byref(p))
        return Tensor(lay, self.dev.UsingPtr(p.value, lay.NElems, dtype, owner=self))

    @staticmethod
    def _d(t: Tensor):
        return t.Backend._d(t)

    # -- reductions ---------------------------------------------------------------------------------------------
    def reduce_axis(self, rank: int, member: str, local: Tensor, axis: int, total_rows: int, row_begin: int,
                    value=None, out: Optional[Tensor] = None) -> Tensor:
        """`member` is the ITensorBackend member name (SumLastAxis, ArgMaxLastAxis, FindLastAxis, ...). `local` is
        rank's slab — rows [row_begin, row_begin + local.Shape[0]) of a tensor whose dim 0 has `total_rows` rows.
        Returns the FULL result (valid on the rank's stream once the call has been issued on every rank)."""
        out_dt = _OUT_DTYPE.get(member, local.DataType)
        if axis != 0:
            full_shape = (total_rows,) + tuple(n for d, n in enumerate(local.Shape) if d not in (0, axis))
            if out is None:
                out = self.alloc(rank, full_shape, out_dt)
            trgt_local_shape = (local.Shape[0],) + full_shape[1:]
            # the frontend's axis -> last permutation (Tensor.fs:4548-4560) on the slab
            probe = Tensor(TL.newC(trgt_local_shape), out.Storage)
            src = Tensor.PrepareAxisReduceSources(probe, axis, local)
            if member == "FindLastAxis":
                keep, p = native.scalar_buffer(value, local.DataType)
                self.api.call("shard_find_last_axis", self._g, rank, p, self._d(out), row_begin, self._d(src))
            elif member in _ARG:
                self.api.call("shard_arg_reduce_last_axis", self._g, rank, _ARG[member], self._d(out), row_begin,
                              self._d(src))
            else:
                self.api.call("shard_reduce_last_axis", self._g, rank, REDUCE_OPS.index(member), self._d(out),
                              row_begin, self._d(src))
            return out
        # the sharded axis itself (incl. whole-tensor folds of a flattened slab)
        shape = tuple(local.Shape[1:])
        if out is None:
            out = self.alloc(rank, shape, out_dt)
        probe = Tensor(TL.newC(shape), out.Storage)
        src = Tensor.PrepareAxisReduceSources(probe, 0, local)
        if member == "FindLastAxis":
            keep, p = native.scalar_buffer(value, local.DataType)
            kind, op = 2, 0
        elif member in _ARG:
            kind, op, p = 1, _ARG[member], None
        else:
            kind, op, p = 0, REDUCE_OPS.index(member), None
        self.api.call("shard_reduce_sharded_axis", self._g, rank, kind, op, p, self._d(out), row_begin, self._d(src))
        return out

    def minmax_arg(self, rank: int, is_max: bool, local: Tensor, total_rows: int, row_begin: int,
                   out_val: Optional[Tensor] = None, out_idx: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Max + ArgMax (or Min + ArgMin) over the LAST axis of a 2-D+ slab in one pass over it."""
        full_shape = (total_rows,) + tuple(local.Shape[1:-1])
        if out_val is None:
            out_val = self.alloc(rank, full_shape, local.DataType)
        if out_idx is None:
            out_idx = self.alloc(rank, full_shape, dtypes.DN_I64)
        self.api.call("shard_minmax_arg_last_axis", self._g, rank, 1 if is_max else 0, self._d(out_val),
                      self._d(out_idx), row_begin, self._d(local))
        return out_val, out_idx

    # -- ordered compaction (row-major order of the full tensor == rank order of the slabs) ------------------------
    def count_true(self, masks: Dict[int, Tensor]) -> List[int]:
        """counts[r] = number of true elements of rank r's mask slab (blocking). `masks`: local rank -> slab."""
        with self.bracket():
            for r in self.ranks:
                self.api.call("shard_count_true_begin", self._g, r, self._d(masks[r]))
        counts = (C.c_int64 * self.world)()
        for r in self.ranks:
            self.api.call("shard_count_true_end", self._g, r, counts)
        return [int(c) for c in counts]

    def true_indices(self, masks: Dict[int, Tensor], row_begins: Dict[int, int]) -> Dict[int, Tensor]:
        """Tensor.trueIdx of a bool tensor sharded along dim 0: {local rank: full [nTrue, nDims] result}."""
        counts = self.count_true(masks)
        total = sum(counts)
        out = {}
        with self.bracket():
            for r in self.ranks:
                m = masks[r]
                out[r] = self.alloc(r, (total, m.NDims), dtypes.DN_I64)
                self.api.call("shard_true_indices", self._g, r, self._d(out[r]), sum(counts[:r]), counts[r], self._d(m),
                              row_begins[r])
        return out

    def masked_get(self, locals_: Dict[int, Tensor], masks: Dict[int, Tensor]) -> Dict[int, Tensor]:
        """`a.M(mask)` with a mask of a's full shape, both sharded along dim 0: {local rank: full 1-D result}."""
        counts = self.count_true(masks)
        total = sum(counts)
        out = {}
        with self.bracket():
            for r in self.ranks:
                a = locals_[r]
                out[r] = self.alloc(r, (total,), a.DataType)
                self.api.call("shard_masked_get", self._g, r, self._d(out[r]), sum(counts[:r]), counts[r], self._d(a),
                              self._d(masks[r]))
        return out
