"""Leading-axis sharding over 2 ranks with the `gloo` backend on CPU (the compute device is the HostTensor oracle,
so this exercises the sharding / ordered-combine ALGORITHM restated in oracle/shard_gloo.py; the product path — the
dn_shard_* entry points over peer memory — is covered by tests/test_shard_gpu.py against the same oracle): sharded results must equal
the unsharded ones bit for bit for integer / index results, and for float Min/Max including the NaN rules."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepnet_b200 import NotFound, Tensor, dtypes
        from deepnet_b200 import layout as TL
        from oracle.shard_gloo import LeadingAxisSharding, slab
        from oracle.host_tensor import HostTensor, TensorHostStorage

        def wrap(t: torch.Tensor) -> Tensor:
            return Tensor(TL.newC(tuple(t.shape)), TensorHostStorage(t.numpy().reshape(-1), HostTensor.Dev))

        sh = LeadingAxisSharding(wrap, torch.device("cpu"))
        rng = np.random.default_rng(7)
        R, C = 101, 37   # odd sizes: ranks get different slab sizes
        f = rng.uniform(-50, 50, size=(R, C)).astype(np.float32)
        f[10, 3] = np.nan          # NaN in rank 0's slab
        f[90, 5] = np.nan          # NaN in rank 1's slab
        f[:, 7] = -np.inf          # column that ArgMax cannot beat -> NotFound
        f[60, 9] = f[:, 9].max() + 1
        f[20, 9] = f[60, 9]        # tie across ranks: lowest global index wins
        i = np.rint(rng.uniform(-50, 50, size=(R, C))).astype(np.int64)
        b = rng.uniform(0, 1, size=(R, C)) < 0.9
        b[:, 4] = True
        full_f, full_i, full_b = HostTensor.ofNumpy(f), HostTensor.ofNumpy(i), HostTensor.ofNumpy(b)
        beg, cnt = slab(R, rank, world)
        loc_f, loc_i, loc_b = (HostTensor.ofNumpy(x[beg:beg + cnt]) for x in (f, i, b))
        checks = []

        def same(name, got: Tensor, want: Tensor, exact=True):
            g, w = got.toNumpy(), want.toNumpy()
            ok = g.shape == w.shape and bool((((g == w) | (np.isnan(g.astype(np.float64)) & np.isnan(w.astype(np.float64))))
                                              if exact else np.isclose(g, w, rtol=1e-5, atol=1e-3, equal_nan=True)).all())
            checks.append((name, ok))

        for member, fn in [("SumLastAxis", "sumAxis"), ("MinLastAxis", "minAxis"), ("MaxLastAxis", "maxAxis"),
                           ("ArgMinLastAxis", "argMinAxis"), ("ArgMaxLastAxis", "argMaxAxis")]:
            for axis in (0, 1):
                same(f"i64 {member} axis {axis}", sh.reduce_axis(member, loc_i, axis, R), getattr(full_i, fn)(axis))
                same(f"f32 {member} axis {axis}", sh.reduce_axis(member, loc_f, axis, R), getattr(full_f, fn)(axis),
                     exact=member != "SumLastAxis")
        for member, fn in [("AllLastAxis", "allAxis"), ("AnyLastAxis", "anyAxis"), ("CountTrueLastAxis", "countTrueAxis")]:
            for axis in (0, 1):
                same(f"bool {member} axis {axis}", sh.reduce_axis(member, loc_b, axis, R), getattr(full_b, fn)(axis))
        for axis in (0, 1):
            same(f"find axis {axis}", sh.reduce_axis("FindLastAxis", loc_i, axis, R, value=7), full_i.findAxis(7, axis))
        # whole-tensor reductions of a flattened shard (1-D, sharded along its only axis)
        flat_full, flat_loc = full_i.flatten(), loc_i.flatten()
        same("whole argmax", sh.reduce_axis("ArgMaxLastAxis", flat_loc, 0, R * C), flat_full.argMaxAxis(0))
        same("whole sum", sh.reduce_axis("SumLastAxis", flat_loc, 0, R * C), flat_full.sumAxis(0))
        # element-wise needs no collective: compute on the slab, gather rows, compare
        same("elementwise + gather rows", sh.all_gather_rows(loc_i * 3 + loc_i, R), full_i * 3 + full_i)
        # equal slabs (100 rows over 2 ranks): the in-place all-gather path of the non-sharded-axis reductions
        R2 = 100
        b2, c2 = slab(R2, rank, world)
        ef, ei, eb = HostTensor.ofNumpy(f[:R2]), HostTensor.ofNumpy(i[:R2]), HostTensor.ofNumpy(b[:R2])
        lf, li, lb = (HostTensor.ofNumpy(x[b2:b2 + c2]) for x in (f, i, b))
        for member, fn in [("SumLastAxis", "sumAxis"), ("MaxLastAxis", "maxAxis"), ("ArgMaxLastAxis", "argMaxAxis"),
                           ("ArgMinLastAxis", "argMinAxis")]:
            same(f"equal slabs i64 {member}", sh.reduce_axis(member, li, 1, R2), getattr(ei, fn)(1))
            same(f"equal slabs f32 {member}", sh.reduce_axis(member, lf, 1, R2), getattr(ef, fn)(1),
                 exact=member != "SumLastAxis")
        same("equal slabs countTrue", sh.reduce_axis("CountTrueLastAxis", lb, 1, R2), eb.countTrueAxis(1))
        same("equal slabs find", sh.reduce_axis("FindLastAxis", li, 1, R2, value=7), ei.findAxis(7, 1))
        # ordered compaction across shards: trueIdx coordinates and MaskedGet values in global row-major order
        same("trueIdx", sh.true_indices(loc_b, R), full_b.trueIdx())
        sparse = rng.uniform(0, 1, size=(R, C)) < 0.05
        sparse[:beg + cnt if rank == 0 else R] &= True
        same("trueIdx sparse", sh.true_indices(HostTensor.ofNumpy(sparse[beg:beg + cnt]), R), HostTensor.ofNumpy(sparse).trueIdx())
        same("maskedGet", sh.masked_get(loc_i, loc_b), full_i.M(full_b))
        none = np.zeros((R, C), dtype=bool)
        none[R - 1, C - 1] = True   # only the last rank selects anything
        same("maskedGet one", sh.masked_get(loc_f, HostTensor.ofNumpy(none[beg:beg + cnt])), full_f.M(HostTensor.ofNumpy(none)))
        bad = [n for n, ok in checks if not ok]
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("OK" if not bad else "FAIL: " + "; ".join(bad))
    finally:
        dist.destroy_process_group()


def test_sharded_reductions_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "OK"


def test_slab_partition():
    from oracle.shard_gloo import slab
    for n in (0, 1, 7, 8, 101, 16384):
        for w in (1, 2, 4, 8):
            parts = [slab(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
