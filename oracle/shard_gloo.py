"""TEST INFRASTRUCTURE — the leading-axis sharding ALGORITHM restated over torch.distributed, so
that the N > 1 logic (slab partition, ordered combine of partials, ragged concatenation) can be checked on CPU with
the `gloo` backend and the HostTensor oracle as the compute device (tests/test_shard_cpu.py). The product path is
deepnet_b200/shard.py over the dn_shard_* entry points of libdeepnet_b200.so (peer-memory stores, no
torch.distributed); nothing in deepnet_b200/ imports this module.

One process per GPU (torchrun); every rank holds a contiguous slab of dim 0 of every operand. Element-wise operators
need no communication. Reductions:
  * over an axis other than 0: every output row lives on one rank -> outputs are all-gathered (KB-MB messages);
  * over axis 0 (the sharded axis, which includes whole-tensor reductions of a flattened shard): every rank
    reduces its slab, the per-rank partials are all-gathered and folded locally in RANK ORDER with the same backend
    operators, so the result is identical on every rank and deterministic. Float Min/Max keep the host's
    order-dependent NaN behaviour (a NaN seen by a later rank resets the running value), ArgMin/ArgMax keep
    first-occurrence semantics through (value, global index) pairs with lowest-index tie-break, Find takes the
    minimum global index.
The bulk tensors never move; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is only plumbing.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from deepnet_b200 import dtypes
from deepnet_b200.layout import NotFound
from deepnet_b200.tensor import Tensor

_TORCH = {dtypes.DN_F32: torch.float32, dtypes.DN_F64: torch.float64, dtypes.DN_I8: torch.int8,
          dtypes.DN_U8: torch.uint8, dtypes.DN_I16: torch.int16, dtypes.DN_I32: torch.int32,
          dtypes.DN_I64: torch.int64, dtypes.DN_BOOL: torch.bool}


def slab(nrows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [begin, begin+count) of `nrows` rows for `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(nrows, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


class LeadingAxisSharding:
    """`wrap(torch_tensor) -> Tensor` exposes a torch buffer on the compute device as a Tensor without copying
    (CudaTensor.usingPtr on GPUs); `torch_device` is where those buffers live."""

    def __init__(self, wrap: Callable[[torch.Tensor], Tensor], torch_device, group=None):
        self.wrap = wrap
        self.torch_device = torch_device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _buffer(self, shape, dtype: int) -> Tuple[torch.Tensor, Tensor]:
        if dtype not in _TORCH:
            raise NotImplementedError(f"sharded collectives for dtype {dtypes.NAMES[dtype]}")
        t = torch.empty(tuple(shape) if len(shape) else (1,), dtype=_TORCH[dtype], device=self.torch_device)
        w = self.wrap(t)
        return t, (w if len(shape) else w.reshape(()))

    def _sync_compute(self, x: Tensor) -> None:
        """Orders the compute device's work before a collective. When the backend launches on torch's current CUDA
        stream (dn_set_stream(torch stream), as bench.py and the tests do) torch.distributed already orders its
        NCCL stream after it with an event, and no host synchronisation is needed."""
        get = getattr(x.Dev, "GetStream", None)
        if get is not None and self.torch_device.type == "cuda":
            if get() == torch.cuda.current_stream(self.torch_device).cuda_stream:
                return
        sync = getattr(x.Dev, "Synchronize", None)
        if sync is not None:
            sync()

    def all_gather_parts(self, local: Tensor) -> Tensor:
        """Every rank contributes one tensor of the same shape; returns [world, *shape] on every rank."""
        send_t, send = self._buffer(local.Shape, local.DataType)
        send.CopyFrom(local)
        recv_t, recv = self._buffer((self.world,) + local.Shape, local.DataType)
        self._sync_compute(local)
        if self.world > 1:
            dist.all_gather_into_tensor(recv_t.view(-1), send_t.view(-1), group=self.group)
        else:
            recv_t.view(-1).copy_(send_t.view(-1))
        return recv

    def all_gather_rows(self, local: Tensor, total_rows: int) -> Tensor:
        """Concatenates the per-rank row slabs (possibly of different sizes) into the full [total_rows, ...]."""
        maxc = slab(total_rows, 0, self.world)[1]
        rest = local.Shape[1:]
        padded_t, padded = self._buffer((maxc,) + rest, local.DataType)
        if local.Shape[0] > 0:
            padded[0:local.Shape[0]].CopyFrom(local)
        parts = self.all_gather_parts(padded)            # [world, maxc, ...]
        _, out = self._buffer((total_rows,) + rest, local.DataType)
        for r in range(self.world):
            b, c = slab(total_rows, r, self.world)
            if c > 0:
                out[b:b + c].CopyFrom(parts[r][0:c])
        return out

    def all_gather_ragged(self, local: Tensor) -> Tensor:
        """Concatenates per-rank tensors whose dim 0 differs from rank to rank (compaction results), in rank order.
        NCCL has no allgatherv: the counts are all-gathered first, then every rank broadcasts its block into its
        place of the result."""
        cnt_t, cnt = self._buffer((1,), dtypes.DN_I64)
        cnt.FillConst(local.Shape[0])
        counts = [int(c) for c in self.all_gather_parts(cnt).toNumpy().reshape(-1)]
        rest = local.Shape[1:]
        out_t, out = self._buffer((sum(counts),) + rest, local.DataType)
        row = 1
        for n in rest:
            row *= n
        off = 0
        for r, c in enumerate(counts):
            if c > 0:
                if r == self.rank:
                    out[off:off + c].CopyFrom(local)
                    self._sync_compute(local)
                if self.world > 1:
                    dist.broadcast(out_t.view(-1)[off * row:(off + c) * row], src=dist.get_global_rank(self.group, r)
                                   if self.group is not None else r, group=self.group)
            off += c
        return out

    # -- ordered compaction (row-major order of the full tensor == rank order of the slabs) -----------------------
    def true_indices(self, local_mask: Tensor, total_rows: int) -> Tensor:
        """Tensor.trueIdx of a bool tensor sharded along dim 0: local TrueIndices, dim-0 coordinates shifted by the
        slab's first row, blocks concatenated in rank order."""
        base, _ = slab(total_rows, self.rank, self.world)
        idx = local_mask.trueIdx()                                   # [nTrue_local, nDims]
        if base != 0 and idx.Shape[0] > 0:
            col0 = idx[:, 0:1]
            col0.CopyFrom(col0 + base)
        return self.all_gather_ragged(idx)

    def masked_get(self, local: Tensor, local_mask: Tensor) -> Tensor:
        """`a.M(mask)` with a mask of a's full shape, both sharded along dim 0: local MaskedGet, blocks concatenated
        in rank order (the logical row-major walk of ScalarOps.fs:667-681 visits the slabs in that order)."""
        return self.all_gather_ragged(local.M(local_mask))

    # -- reductions -------------------------------------------------------------------------------------------
    # result dtype of each member when it differs from the source's (None = same as the source)
    OUT_DTYPE = {"SumLastAxis": None, "ProductLastAxis": None, "MinLastAxis": None, "MaxLastAxis": None,
                 "AllLastAxis": None, "AnyLastAxis": None, "CountTrueLastAxis": dtypes.DN_I64,
                 "ArgMinLastAxis": dtypes.DN_I64, "ArgMaxLastAxis": dtypes.DN_I64, "FindLastAxis": dtypes.DN_I64}
    FOLDS = {"SumLastAxis": "sumAxis", "ProductLastAxis": "productAxis", "MinLastAxis": "minAxis",
             "MaxLastAxis": "maxAxis", "AllLastAxis": "allAxis", "AnyLastAxis": "anyAxis"}

    def reduce_axis(self, member: str, local: Tensor, axis: int, total_rows: int, value=None) -> Tensor:
        """`member` is the ITensorBackend member name (SumLastAxis, ArgMaxLastAxis, FindLastAxis, ...). `local` is
        this rank's slab of a tensor whose dim 0 has `total_rows` rows in total. Returns the full result."""
        local_fn = {**self.FOLDS, "CountTrueLastAxis": "countTrueAxis", "ArgMinLastAxis": "argMinAxis",
                    "ArgMaxLastAxis": "argMaxAxis"}
        if axis != 0:
            if total_rows % self.world == 0 and local.Shape[0] == total_rows // self.world and member in self.OUT_DTYPE:
                # equal slabs: every rank reduces straight into its rows of the full result, then ONE in-place
                # all-gather (send buffer = this rank's block of the receive buffer) combines them
                out_dt = self.OUT_DTYPE[member] if self.OUT_DTYPE[member] is not None else local.DataType
                rest = tuple(n for d, n in enumerate(local.Shape) if d != axis)[1:]
                out_t, out = self._buffer((total_rows,) + rest, out_dt)
                cnt = local.Shape[0]
                mine = out[self.rank * cnt:(self.rank + 1) * cnt]
                if member == "FindLastAxis":
                    src = Tensor.PrepareAxisReduceSources(mine, axis, local)
                    src.Backend.FindLastAxis(value, mine, src)
                else:
                    mine._fill_axis(member, axis, local, on_src_backend=self.OUT_DTYPE[member] is not None)
                self._sync_compute(local)
                if self.world > 1:
                    flat = out_t.view(-1)
                    per = flat.numel() // self.world
                    dist.all_gather_into_tensor(flat, flat[self.rank * per:(self.rank + 1) * per], group=self.group)
                return out
            part = local.findAxis(value, axis) if member == "FindLastAxis" else getattr(local, local_fn[member])(axis)
            return self.all_gather_rows(part, total_rows)
        base, _ = slab(total_rows, self.rank, self.world)
        if member in ("SumLastAxis", "ProductLastAxis", "AllLastAxis", "AnyLastAxis", "CountTrueLastAxis"):
            parts = self.all_gather_parts(getattr(local, local_fn[member])(0))
            combine = "sumAxis" if member == "CountTrueLastAxis" else local_fn[member]
            return getattr(parts, combine)(0)
        if member in ("MinLastAxis", "MaxLastAxis"):
            return self._minmax_over_shards(member, local)
        if member in ("ArgMinLastAxis", "ArgMaxLastAxis"):
            return self._arg_over_shards(member, local, base)
        if member == "FindLastAxis":
            idx = local.findAxis(value, 0)
            big = Tensor.filled(idx.Shape, 2 ** 62, dtypes.DN_I64, idx.Dev)
            glob = Tensor.ifThenElse(idx.eq(NotFound), big, idx + base)
            best = self.all_gather_parts(glob).minAxis(0)
            return Tensor.ifThenElse(best.eq(2 ** 62), Tensor.filled(best.Shape, NotFound, dtypes.DN_I64, best.Dev), best)
        raise ValueError(member)

    def _minmax_over_shards(self, member: str, local: Tensor) -> Tensor:
        is_max = member == "MaxLastAxis"
        val = local.maxAxis(0) if is_max else local.minAxis(0)
        parts = self.all_gather_parts(val)
        if not dtypes.is_float(local.DataType):
            return parts.maxAxis(0) if is_max else parts.minAxis(0)
        # ordered fold with the host's NaN rule (ScalarOps.fs:620-628): a shard that contains a NaN replaces the
        # running value with its own result; otherwise res = if res `better` v then res else v
        has_nan = self.all_gather_parts(local.ne(local).anyAxis(0))
        acc = parts[0]
        for r in range(1, self.world):
            v = parts[r]
            better = acc.gt(v) if is_max else acc.lt(v)
            folded = Tensor.ifThenElse(better, acc, v)
            acc = Tensor.ifThenElse(has_nan[r], v, folded)
        return acc.Copy()

    def _arg_over_shards(self, member: str, local: Tensor, base: int) -> Tensor:
        is_max = member == "ArgMaxLastAxis"
        idx = local.argMaxAxis(0) if is_max else local.argMinAxis(0)
        found = idx.ne(NotFound)
        safe = Tensor.ifThenElse(found, idx, Tensor.zeros(idx.Shape, dtypes.DN_I64, idx.Dev))
        # value at the arg position: gather along dim 0; the remaining source dims d >= 1 follow target dim d-1
        # (gather's `None` means "same dim number", which is off by one here, so the identity maps are explicit)
        shp = idx.Shape
        ident = []
        for j, n in enumerate(shp):
            cnt = Tensor.counting(idx.Dev, n).reshape([1] * j + [n] + [1] * (len(shp) - j - 1))
            ident.append(cnt.broadcastTo(shp))
        val = Tensor.gather([safe] + ident, local) if local.Shape[0] > 0 else \
            Tensor.zeros(idx.Shape, local.DataType, idx.Dev)
        glob = Tensor.ifThenElse(found, idx + base, idx)
        vals, idxs = self.all_gather_parts(val), self.all_gather_parts(glob)
        native = getattr(vals.Backend.api, "_arg_reduce_combine", None)
        if native is not None and vals.NDims == 2:
            out = Tensor.empty(idx.Shape, dtypes.DN_I64, idx.Dev)
            be = vals.Backend
            vals.Backend.api.call("arg_reduce_combine", 1 if is_max else 0, be._d(out), be._d(vals), be._d(idxs))
            return out
        acc_v, acc_i = vals[0], idxs[0]
        for r in range(1, self.world):
            v, i = vals[r], idxs[r]
            better = v.gt(acc_v) if is_max else v.lt(acc_v)
            take = i.ne(NotFound) & (acc_i.eq(NotFound) | better)   # strict compare keeps the lowest global index
            acc_v = Tensor.ifThenElse(take, v, acc_v)
            acc_i = Tensor.ifThenElse(take, i, acc_i)
        return acc_i.Copy()
