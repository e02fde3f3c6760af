"""The sharded-operator parity cases shared by the single-process and the multi-process tests of the dn_shard_* path
(tests/test_shard_gpu.py): every sharded result is compared with the UNSHARDED HostTensor oracle on the same seeded
data (host semantics: ScalarOps.fs:606-707). Bit-exact for integer / bool / index results and for float Min/Max
(NaN rules included); float Sum within rel 1e-4*log2(n)."""
from __future__ import annotations

import os

import numpy as np

from deepnet_b200 import CudaTensor, dtypes
from deepnet_b200.shard import ShardGroup, slab
from oracle.host_tensor import HostTensor


def _same(name, got: np.ndarray, want: np.ndarray, bad, rtol=0.0):
    ok = got.shape == want.shape and got.dtype == want.dtype
    if ok:
        g, w = got.astype(np.float64), want.astype(np.float64)
        if rtol:
            ok = bool(np.allclose(g, w, rtol=rtol, atol=0.5, equal_nan=True))
        else:
            ok = bool(((got == want) | (np.isnan(g) & np.isnan(w))).all())
    if not ok:
        bad.append(name)


def make_data(R=4099, C=1000):
    rng = np.random.default_rng(3)
    f = rng.uniform(-50, 50, size=(R, C)).astype(np.float32)
    f[5, 1] = np.nan                      # NaN in the first slab
    f[R - 99, 2] = np.nan                 # NaN in the last slab
    f[:, 3] = -np.inf                     # ArgMax cannot beat the initial value -> NotFound
    f[100, 4] = 99.0
    f[R - 1000, 4] = 99.0                 # tie across slabs: lowest global index wins
    f[7] = np.nan                         # an all-NaN row
    f[R - 1, 5] = np.nan                  # column that ENDS with a NaN (Max over dim 0 stays NaN)
    i = np.rint(rng.uniform(-50, 50, size=(R, C))).astype(np.int64)
    b = rng.uniform(0, 1, size=(R, C)) < 0.9
    b[:, 4] = True
    sparse = rng.uniform(0, 1, size=(R, C)) < 0.02
    return f, i, b, sparse


def run(grp: ShardGroup, set_device) -> list:
    """Runs every case on the group's local ranks. `set_device(rank)` binds the calling thread to the rank's device
    for the ordinary (non-collective) calls. Returns the names of the failed checks."""
    W = grp.world
    f, i, b, sparse = make_data()
    R, C = f.shape
    hf, hi, hb, hs = (HostTensor.ofNumpy(x) for x in (f, i, b, sparse))
    loc = {}
    for r in grp.ranks:
        set_device(r)
        beg, cnt = slab(R, r, W)
        loc[r] = dict(beg=beg, f=CudaTensor.ofNumpy(f[beg:beg + cnt]), i=CudaTensor.ofNumpy(i[beg:beg + cnt]),
                      b=CudaTensor.ofNumpy(b[beg:beg + cnt]), s=CudaTensor.ofNumpy(sparse[beg:beg + cnt]))
        grp.dev.Synchronize()        # uploads ran on the thread's stream; the collectives run on the rank's stream
    bad = []
    results = []   # (name, {rank: Tensor}, oracle ndarray, rtol)

    def collect(name, fn, want, rtol=0.0):
        outs = {}
        with grp.bracket():
            for r in grp.ranks:
                set_device(r)
                outs[r] = fn(r)
        if os.environ.get("DN_SHARD_DEBUG"):
            grp.sync()
            print("done:", name, flush=True)
        results.append((name, outs, want.toNumpy() if hasattr(want, "toNumpy") else want, rtol))

    members = [("SumLastAxis", "sumAxis"), ("ProductLastAxis", "productAxis"), ("MinLastAxis", "minAxis"),
               ("MaxLastAxis", "maxAxis"), ("ArgMinLastAxis", "argMinAxis"), ("ArgMaxLastAxis", "argMaxAxis")]
    for member, fn in members:
        for axis in (0, 1):
            if member != "ProductLastAxis":
                collect(f"f32 {member} axis {axis}",
                        lambda r, m=member, a=axis: grp.reduce_axis(r, m, loc[r]["f"], a, R, loc[r]["beg"]),
                        getattr(hf, fn)(axis), rtol=1.4e-3 if member == "SumLastAxis" else 0.0)
            collect(f"i64 {member} axis {axis}",
                    lambda r, m=member, a=axis: grp.reduce_axis(r, m, loc[r]["i"], a, R, loc[r]["beg"]),
                    getattr(hi, fn)(axis))
    for member, fn in [("AllLastAxis", "allAxis"), ("AnyLastAxis", "anyAxis"), ("CountTrueLastAxis", "countTrueAxis")]:
        for axis in (0, 1):
            collect(f"bool {member} axis {axis}",
                    lambda r, m=member, a=axis: grp.reduce_axis(r, m, loc[r]["b"], a, R, loc[r]["beg"]),
                    getattr(hb, fn)(axis))
    for axis in (0, 1):
        collect(f"find axis {axis}",
                lambda r, a=axis: grp.reduce_axis(r, "FindLastAxis", loc[r]["i"], a, R, loc[r]["beg"], value=7),
                hi.findAxis(7, axis))
        collect(f"find f32 99 axis {axis}",
                lambda r, a=axis: grp.reduce_axis(r, "FindLastAxis", loc[r]["f"], a, R, loc[r]["beg"], value=99.0),
                hf.findAxis(99.0, axis))
    # whole-tensor folds of the flattened slab (1-D, sharded along its only axis)
    for member, fn in [("ArgMaxLastAxis", "argMaxAxis"), ("SumLastAxis", "sumAxis"), ("MaxLastAxis", "maxAxis")]:
        collect(f"whole i64 {member}",
                lambda r, m=member: grp.reduce_axis(r, m, loc[r]["i"].flatten(), 0, R * C, loc[r]["beg"] * C),
                getattr(hi.flatten(), fn)(0))
    collect("whole f32 ArgMinLastAxis",
            lambda r: grp.reduce_axis(r, "ArgMinLastAxis", loc[r]["f"].flatten(), 0, R * C, loc[r]["beg"] * C),
            hf.flatten().argMinAxis(0))
    collect("whole f32 MaxLastAxis (NaN rule)",
            lambda r: grp.reduce_axis(r, "MaxLastAxis", loc[r]["f"].flatten(), 0, R * C, loc[r]["beg"] * C),
            hf.flatten().maxAxis(0))
    # a sliced, reversed slab view (non-contiguous source)
    collect("f32 MaxLastAxis of a reversed slice",
            lambda r: grp.reduce_axis(r, "MaxLastAxis", loc[r]["f"][:, 10:].reverseAxis(1), 1, R, loc[r]["beg"]),
            hf[:, 10:].reverseAxis(1).maxAxis(1))
    # the same target twice in a row: one rank per process -> the library inserts the entry barrier itself; one
    # thread driving several ranks -> an explicit barrier bracket in between (the library refuses otherwise)
    reuse = {}
    for rep in range(2):
        if rep and len(grp.ranks) > 1:
            grp.barrier()
        with grp.bracket():
            for r in grp.ranks:
                set_device(r)
                if r not in reuse:
                    reuse[r] = grp.alloc(r, (R,), dtypes.DN_I64)
                grp.reduce_axis(r, "ArgMaxLastAxis" if rep else "ArgMinLastAxis", loc[r]["i"], 1, R, loc[r]["beg"],
                                out=reuse[r])
    results.append(("same target twice", reuse, hi.argMaxAxis(1).toNumpy(), 0.0))
    # Max + ArgMax in one pass
    fused_v, fused_i = {}, {}
    with grp.bracket():
        for r in grp.ranks:
            set_device(r)
            fused_v[r], fused_i[r] = grp.minmax_arg(r, True, loc[r]["f"], R, loc[r]["beg"])
    results.append(("fused Max", fused_v, hf.maxAxis(1).toNumpy(), 0.0))
    results.append(("fused ArgMax", fused_i, hf.argMaxAxis(1).toNumpy(), 0.0))
    fused_v2, fused_i2 = {}, {}
    with grp.bracket():
        for r in grp.ranks:
            set_device(r)
            fused_v2[r], fused_i2[r] = grp.minmax_arg(r, False, loc[r]["f"], R, loc[r]["beg"])
    results.append(("fused Min", fused_v2, hf.minAxis(1).toNumpy(), 0.0))
    results.append(("fused ArgMin", fused_i2, hf.argMinAxis(1).toNumpy(), 0.0))
    # several collectives per rank in ONE bracket (one wait at the end): ArgMax, Max and a sharded-axis Sum (last)
    bi, bm, bs = {}, {}, {}
    with grp.batch():
        for r in grp.ranks:
            set_device(r)
            bi[r] = grp.reduce_axis(r, "ArgMaxLastAxis", loc[r]["f"], 1, R, loc[r]["beg"])
            bm[r] = grp.reduce_axis(r, "MaxLastAxis", loc[r]["f"], 1, R, loc[r]["beg"])
            bs[r] = grp.reduce_axis(r, "SumLastAxis", loc[r]["i"], 0, R, loc[r]["beg"])
    results.append(("bracket of three: ArgMax", bi, hf.argMaxAxis(1).toNumpy(), 0.0))
    results.append(("bracket of three: Max", bm, hf.maxAxis(1).toNumpy(), 0.0))
    results.append(("bracket of three: Sum over the sharded axis", bs, hi.sumAxis(0).toNumpy(), 0.0))
    # ordered compaction across the shards
    begs = {r: loc[r]["beg"] for r in grp.ranks}
    results.append(("trueIdx", grp.true_indices({r: loc[r]["b"] for r in grp.ranks}, begs), hb.trueIdx().toNumpy(), 0.0))
    results.append(("trueIdx sparse", grp.true_indices({r: loc[r]["s"] for r in grp.ranks}, begs), hs.trueIdx().toNumpy(), 0.0))
    results.append(("maskedGet i64", grp.masked_get({r: loc[r]["i"] for r in grp.ranks}, {r: loc[r]["b"] for r in grp.ranks}),
                    hi.M(hb).toNumpy(), 0.0))
    results.append(("maskedGet f32 sparse", grp.masked_get({r: loc[r]["f"] for r in grp.ranks}, {r: loc[r]["s"] for r in grp.ranks}),
                    hf.M(hs).toNumpy(), 0.0))
    grp.sync()
    for name, outs, want, rtol in results:
        for r, t in outs.items():
            set_device(r)
            _same(f"{name} (rank {r})", t.toNumpy(), want, bad, rtol)
    return bad
