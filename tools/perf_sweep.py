"""Per-op bandwidth sweep on one GPU (development tool; bench.py is the contract harness).
Usage: python tools/perf_sweep.py [--n 28] [--reps 5] [--filter substr]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from deepnet_b200 import CudaTensor, Tensor, dtypes

TORCH_DT = {dtypes.DN_F32: torch.float32, dtypes.DN_F64: torch.float64, dtypes.DN_I32: torch.int32,
            dtypes.DN_I64: torch.int64, dtypes.DN_BOOL: torch.bool, dtypes.DN_U8: torch.uint8}


def wrap(t: torch.Tensor) -> Tensor:
    dt = {v: k for k, v in TORCH_DT.items()}[t.dtype]
    return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)


def rand(shape, dt):
    if dt in (dtypes.DN_F32, dtypes.DN_F64):
        return (torch.rand(shape, device="cuda", dtype=TORCH_DT[dt]) * 100 - 50)
    if dt == dtypes.DN_BOOL:
        return torch.rand(shape, device="cuda") >= 0.5
    return torch.randint(-50, 50, shape, device="cuda", dtype=TORCH_DT[dt])


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--filter", default="")
    args = ap.parse_args()
    dev = CudaTensor.dev()
    dev.Init(0)
    dev.SetStream(torch.cuda.current_stream().cuda_stream)
    side = 1 << (args.n // 2)
    R, Cc = side, (1 << args.n) // side
    peak = 6539.9
    rows = []

    def run(name, nbytes, fn):
        if args.filter and args.filter not in name:
            return
        ms = timeit(fn, args.reps)
        gbs = nbytes / ms / 1e6
        rows.append((name, ms, gbs))
        print(f"{name:48s} {ms:9.3f} ms {gbs:9.1f} GB/s  {100 * gbs / peak:5.1f}% of measured copy peak", flush=True)

    # torch copy ceiling on this box, same size
    x = torch.empty(1 << args.n, device="cuda", dtype=torch.float32)
    y = torch.empty_like(x)
    run("torch copy_ f32 (ceiling)", 2 * x.numel() * 4, lambda: y.copy_(x))
    del x, y

    for dt, s in ((dtypes.DN_F32, 4), (dtypes.DN_F64, 8), (dtypes.DN_I32, 4)):
        nm = dtypes.NAMES[dt]
        N = R * Cc
        ta, tb, tc = rand((R, Cc), dt), rand((R, Cc), dt), torch.empty((R, Cc), device="cuda", dtype=TORCH_DT[dt])
        a, b, c = wrap(ta), wrap(tb), wrap(tc)
        trow, tcol = rand((1, Cc), dt), rand((R, 1), dt)
        row, col = wrap(trow), wrap(tcol)
        run(f"C2 {nm} add contiguous", 3 * N * s, lambda: c.FillAdd(a, b))
        run(f"C2 {nm} add a.T + b", 3 * N * s, lambda: c.FillAdd(a.T, b))
        run(f"C2 {nm} add a + row", (2 * N + Cc) * s, lambda: c.FillAdd(a, row))
        run(f"C2 {nm} add a + col", (2 * N + R) * s, lambda: c.FillAdd(a, col))
        if dt != dtypes.DN_I32:
            run(f"C2 {nm} sin(a.T)", 2 * N * s, lambda: c.FillSin(a.T))
            run(f"C2 {nm} sin(a)", 2 * N * s, lambda: c.FillSin(a))
        else:
            run(f"C2 {nm} abs(a.T)", 2 * N * s, lambda: c.FillAbs(a.T))
        cs = c[1:, 1:]
        run(f"C2 {nm} add sliced a[1:,1:] + b[1:,1:]", 3 * (R - 1) * (Cc - 1) * s,
            lambda: cs.FillAdd(a[1:, 1:], b[1:, 1:]))
        run(f"C2 {nm} add reverseAxis0(a) + b", 3 * N * s, lambda: c.FillAdd(a.reverseAxis(0), b))
        run(f"C2 {nm} add reverseAxis1(a) + b", 3 * N * s, lambda: c.FillAdd(a.reverseAxis(1), b))
        run(f"C2 {nm} copy", 2 * N * s, lambda: c.CopyFrom(a))
        run(f"C2 {nm} copy a.T", 2 * N * s, lambda: c.CopyFrom(a.T))
        m = wrap(torch.empty((R, Cc), device="cuda", dtype=torch.bool))
        run(f"C2 {nm} less a < b", (2 * s + 1) * N, lambda: m.FillLess(a, b))
        run(f"C2 {nm} ifThenElse", (3 * s + 1) * N, lambda: c.FillIfThenElse(m, a, b))
        run(f"C2 {nm} fill", N * s, lambda: c.FillConst(3))
        # reductions over the same tensor
        t1 = Tensor.empty((R,), dt, dev)
        t0 = Tensor.empty((Cc,), dt, dev)
        ti = Tensor.empty((R,), dtypes.DN_I64, dev)
        run(f"   {nm} sumAxis 1 [{R},{Cc}]", N * s + R * s, lambda: t1.FillSumAxis(1, a))
        run(f"   {nm} sumAxis 0 [{R},{Cc}]", N * s + Cc * s, lambda: t0.FillSumAxis(0, a))
        run(f"   {nm} maxAxis 1", N * s + R * s, lambda: t1.FillMaxAxis(1, a))
        run(f"   {nm} argMaxAxis 1", N * s + R * 8, lambda: ti._fill_axis("ArgMaxLastAxis", 1, a, True))
        flat = a.reshape((N,))
        ts = Tensor.empty((), dt, dev)
        run(f"   {nm} sum (whole tensor)", N * s, lambda: ts.FillSumAxis(0, flat))
        del ta, tb, tc, a, b, c, m

    # C1
    ta, tb = rand((4096, 4096), dtypes.DN_F32), rand((4096, 4096), dtypes.DN_F32)
    a, b = wrap(ta), wrap(tb)
    t1, t2, c = (Tensor.empty((4096, 4096), dtypes.DN_F32, dev) for _ in range(3))
    out = Tensor.empty((4096,), dtypes.DN_F32, dev)
    n1 = 4096 * 4096 * 4
    run("C1 f32 mul 4096^2", 3 * n1, lambda: t1.FillMultiply(a, b))
    run("C1 f32 sin 4096^2", 2 * n1, lambda: t2.FillSin(a))
    run("C1 f32 add 4096^2", 3 * n1, lambda: c.FillAdd(t1, t2))
    run("C1 f32 sumAxis1 4096^2", n1 + 4096 * 4, lambda: out.FillSumAxis(1, c))
    run("C1 f32 FUSED a*b+sin(a) 4096^2 (one call)", 3 * n1, lambda: c.FillFused(lambda x, y: x * y + x.sin(), a, b))
    # fused expressions at 2^28 elements
    tfa, tfb = rand((R, Cc), dtypes.DN_F32), rand((R, Cc), dtypes.DN_F32)
    fa, fb = wrap(tfa), wrap(tfb)
    fc = wrap(torch.empty((R, Cc), device="cuda", dtype=torch.float32))
    nf = R * Cc * 4
    run("FUSED f32 a*b+sin(a) 2^28", 3 * nf, lambda: fc.FillFused(lambda x, y: x * y + x.sin(), fa, fb))
    run("FUSED f32 dh*(1-h*h) 2^28", 3 * nf, lambda: fc.FillFused(lambda x, y: x * (1.0 - y * y), fa, fb))
    run("FUSED f32 w-g*0.01 in place 2^28", 3 * nf, lambda: fa.FillFused(lambda x, y: x - y * 0.01, fa, fb))
    run("FUSED f32 (a-b)*c/(|c|+2.5) 2^28", 4 * nf, lambda: fc.FillFused(lambda x, y, z: (x - y) * z / (abs(z) + 2.5), fa, fb, fc))
    tda, tdb = rand((R, Cc), dtypes.DN_F64), rand((R, Cc), dtypes.DN_F64)
    da, db = wrap(tda), wrap(tdb)
    dcc = wrap(torch.empty((R, Cc), device="cuda", dtype=torch.float64))
    run("FUSED f64 dh*(1-h*h) 2^28", 3 * nf * 2, lambda: dcc.FillFused(lambda x, y: x * (1.0 - y * y), da, db))
    del tfa, tfb, fa, fb, fc, tda, tdb, da, db, dcc

    # C3
    tl = rand((262144, 1000), dtypes.DN_F32)
    lg = wrap(tl)
    oi = Tensor.empty((262144,), dtypes.DN_I64, dev)
    om = Tensor.empty((262144,), dtypes.DN_F32, dev)
    nb = 262144 * 1000 * 4
    run("C3 f32 argMaxAxis1 262144x1000", nb + 262144 * 8, lambda: oi._fill_axis("ArgMaxLastAxis", 1, lg, True))
    run("C3 f32 maxAxis1 262144x1000", nb + 262144 * 4, lambda: om.FillMaxAxis(1, lg))
    run("C3 f32 sumAxis1 262144x1000", nb + 262144 * 4, lambda: om.FillSumAxis(1, lg))
    _d = lambda t: t.Backend._d(t)
    run("C3 f32 max+argmax ONE pass 262144x1000", nb + 262144 * 12,
        lambda: dev.api.call("shard_minmax_arg_last_axis", None, 0, 1, _d(om), _d(oi), 0, _d(lg)))
    del tl, lg

    # C4
    N = 1 << 26
    tsrc = torch.randint(-(1 << 40), 1 << 40, (N,), device="cuda", dtype=torch.int64)
    tidx = torch.randint(0, N, (N,), device="cuda", dtype=torch.int64)
    tperm = torch.randperm(N, device="cuda")
    src, idx, perm = wrap(tsrc), wrap(tidx), wrap(tperm)
    trg = Tensor.empty((N,), dtypes.DN_I64, dev)
    run("C4 i64 gather random 2^26", 3 * 8 * N, lambda: trg.FillGather([idx], src))
    run("C4 i64 gather permutation 2^26", 3 * 8 * N, lambda: trg.FillGather([perm], src))
    run("C4 i64 scatter random 2^26", (8 + 8 + 8 + 16) * N, lambda: trg.FillScatter([idx], src))
    run("C4 i64 scatter permutation 2^26", (8 + 8 + 8 + 16) * N, lambda: trg.FillScatter([perm], src))
    # C4 variants: hot-spot indices (collisions), 2-D row gather [Some i0; None], sparse masks
    thot = (torch.rand(N, device="cuda") ** 8 * 4096).to(torch.int64)          # most indices fall on a few cells
    hot = wrap(thot)
    run("C4 i64 scatter hot-spot (4096 cells) 2^26", (8 + 8 + 16) * N, lambda: trg.FillScatter([hot], src))
    run("C4 i64 gather hot-spot 2^26", 3 * 8 * N, lambda: trg.FillGather([hot], src))
    tsrc2 = tsrc.view(8192, 8192)
    src2 = wrap(tsrc2)
    trow = torch.randint(0, 8192, (8192, 1), device="cuda", dtype=torch.int64)
    rowidx = wrap(trow).broadcastTo((8192, 8192))
    trg2 = Tensor.empty((8192, 8192), dtypes.DN_I64, dev)
    run("C4 i64 gather rows [Some i0; None] 8192^2", 2 * 8 * N + 8 * 8192, lambda: trg2.FillGather([rowidx, None], src2))
    tms = torch.rand(N, device="cuda") < 0.01
    msp = wrap(tms)
    nsp = int(tms.sum().item())
    gsp = Tensor.empty((nsp,), dtypes.DN_I64, dev)
    run("C4 i64 maskedGet p=0.01 2^26", N + 8 * N + 8 * nsp, lambda: src.Backend.MaskedGet(gsp, src, [msp]))
    run("C4 i64 maskedSet p=0.01 2^26", N + 16 * nsp, lambda: trg.Backend.MaskedSet(trg, [msp], gsp))
    tisp = Tensor.empty((nsp, 2), dtypes.DN_I64, dev)
    run("C4 trueIdx [8192,8192] p=0.01", N + 16 * nsp, lambda: tisp.Backend.TrueIndices(tisp, msp.reshape((8192, 8192))))
    tmask = torch.rand(N, device="cuda") < 0.5
    mask = wrap(tmask)
    ntrue = int(tmask.sum().item())
    got = Tensor.empty((ntrue,), dtypes.DN_I64, dev)
    run("C4 countTrue 2^26", N, lambda: mask.countTrue())
    run("C4 i64 maskedGet p=0.5 2^26", N + 8 * N + 8 * ntrue, lambda: src.Backend.MaskedGet(got, src, [mask]))
    run("C4 i64 maskedSet p=0.5 2^26", N + 8 * ntrue + 8 * ntrue, lambda: trg.Backend.MaskedSet(trg, [mask], got))
    m2 = mask.reshape((8192, 8192))
    ti = Tensor.empty((ntrue, 2), dtypes.DN_I64, dev)
    run("C4 trueIdx [8192,8192] p=0.5", N + 16 * ntrue, lambda: ti.Backend.TrueIndices(ti, m2))
    # A14: VecVecDot / MatVecDot (HBM-bound)
    tv1, tv2 = torch.randn(1 << 28, device="cuda"), torch.randn(1 << 28, device="cuda")
    v1, v2 = wrap(tv1), wrap(tv2)
    sc = Tensor.empty((), dtypes.DN_F32, dev)
    run("A14 f32 VecVecDot 2^28", 2 * 4 * (1 << 28), lambda: sc.FillDot(v1, v2))
    tmat, tx = torch.randn(16384, 16384, device="cuda"), torch.randn(16384, device="cuda")
    mat, xv = wrap(tmat), wrap(tx)
    yv = Tensor.empty((16384,), dtypes.DN_F32, dev)
    run("A14 f32 MatVecDot [16384,16384] . [16384]", 4 * (16384 * 16384 + 2 * 16384), lambda: yv.FillDot(mat, xv))
    run("A14 f32 MatVecDot A.T . x", 4 * (16384 * 16384 + 2 * 16384), lambda: yv.FillDot(mat.T, xv))
    print("launches:", dev.LaunchCount())


if __name__ == "__main__":
    main()
