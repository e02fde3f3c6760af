"""Every ITensorBackend element-wise / reduction member (TensorBackend.fs:67-135) on contiguous >=256 MB tensors:
achieved algorithmic GB/s per op and dtype, against the measured copy peak. Development / evidence tool — the
table it prints is committed under profiles/. Usage: python tools/op_sweep.py [--n 28] [--reps 5] [--out file]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from deepnet_b200 import CudaTensor, Tensor, dtypes
from perf_sweep import TORCH_DT, wrap, timeit

UNARY_ALL = ["UnaryPlus", "UnaryMinus", "Abs", "Sgn"]
UNARY_FLOAT = ["Log", "Log10", "Exp", "Sin", "Cos", "Tan", "Asin", "Acos", "Atan", "Sinh", "Cosh", "Tanh", "Sqrt",
               "Ceiling", "Floor", "Round", "Truncate"]
BINARY = ["Add", "Subtract", "Multiply", "Divide", "Modulo", "Power", "MaxElemwise", "MinElemwise"]
COMPARE = ["Equal", "NotEqual", "Less", "LessOrEqual", "Greater", "GreaterOrEqual"]
REDUCE = ["SumLastAxis", "ProductLastAxis", "MinLastAxis", "MaxLastAxis"]


def domain(member, shape, dt):
    """Inputs inside every function's domain so that no op degenerates into a NaN fast path."""
    tdt = TORCH_DT[dt]
    if dt in (dtypes.DN_F32, dtypes.DN_F64):
        u = torch.rand(shape, device="cuda", dtype=tdt)
        if member in ("Asin", "Acos"):
            return u * 2 - 1
        if member in ("Log", "Log10", "Sqrt", "Power"):
            return u * 50 + 0.01
        if member in ("Exp", "Sinh", "Cosh"):
            return u * 20 - 10
        return u * 100 - 50
    if dt == dtypes.DN_BOOL:
        return torch.rand(shape, device="cuda") >= 0.5
    lo, hi = (1, 50) if member in ("Divide", "Modulo") else (-50, 50)
    return torch.randint(lo, hi, shape, device="cuda", dtype=tdt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--peak", type=float, default=6539.9)
    ap.add_argument("--filter", default="", help="run only the lines whose name contains this substring")
    args = ap.parse_args()
    dev = CudaTensor.dev()
    dev.Init(0)
    dev.SetStream(torch.cuda.current_stream().cuda_stream)
    N = 1 << args.n
    R = 1 << (args.n // 2)
    C = N // R
    lines = []

    def run(name, nbytes, fn):
        if args.filter and args.filter not in name:
            return
        ms = timeit(fn, args.reps)
        gbs = nbytes / ms / 1e6
        line = f"{name:44s} {nbytes / 1e6:9.1f} MB {ms:8.3f} ms {gbs:8.1f} GB/s {100 * gbs / args.peak:6.1f} %"
        lines.append(line)
        print(line, flush=True)

    hdr = f"# op sweep, contiguous [{R},{C}] = 2^{args.n} elements; peak = {args.peak} GB/s (MEASURED_PEAKS.json hbm_gbs)"
    lines.append(hdr)
    print(hdr)
    for dt, s in ((dtypes.DN_F32, 4), (dtypes.DN_F64, 8), (dtypes.DN_I32, 4), (dtypes.DN_I64, 8)):
        nm = dtypes.NAMES[dt]
        isf = dt in (dtypes.DN_F32, dtypes.DN_F64)
        c = wrap(torch.empty((R, C), device="cuda", dtype=TORCH_DT[dt]))
        m = wrap(torch.empty((R, C), device="cuda", dtype=torch.bool))
        for member in UNARY_ALL + (UNARY_FLOAT if isf else []):
            ta = domain(member, (R, C), dt)
            a = wrap(ta)
            run(f"{nm} {member}", 2 * N * s, lambda: getattr(c, "Fill" + member)(a))
            del a, ta
        for member in BINARY:
            if member == "Power" and not isf:
                continue
            ta, tb = domain(member, (R, C), dt), domain(member, (R, C), dt)
            a, b = wrap(ta), wrap(tb)
            run(f"{nm} {member}", 3 * N * s, lambda: getattr(c, "Fill" + member)(a, b))
            del a, b, ta, tb
        ta, tb = domain("Add", (R, C), dt), domain("Add", (R, C), dt)
        a, b = wrap(ta), wrap(tb)
        for member in COMPARE:
            run(f"{nm} {member} -> bool", (2 * s + 1) * N, lambda: getattr(m, "Fill" + member)(a, b))
        if isf:
            run(f"{nm} IsFinite -> bool", (s + 1) * N, lambda: a.Backend.IsFinite(m, a))
        run(f"{nm} IfThenElse", (3 * s + 1) * N, lambda: c.FillIfThenElse(m, a, b))
        run(f"{nm} FillConst", s * N, lambda: c.FillConst(3))
        run(f"{nm} FillIncrementing", s * N, lambda: c.reshape((N,)).FillIncrementing(1, 2))
        run(f"{nm} Copy", 2 * s * N, lambda: c.CopyFrom(a))
        for dt2 in (dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I32, dtypes.DN_I64, dtypes.DN_U8):
            if dt2 == dt:
                continue
            s2 = dtypes.itemsize(dt2)
            c2 = wrap(torch.empty((R, C), device="cuda", dtype=TORCH_DT[dt2]))
            run(f"{nm} Convert -> {dtypes.NAMES[dt2]}", (s + s2) * N, lambda: c2.FillConvert(a))
            del c2
        t1 = Tensor.empty((R,), dt, dev)
        t0 = Tensor.empty((C,), dt, dev)
        ti = Tensor.empty((R,), dtypes.DN_I64, dev)
        for member in REDUCE:
            run(f"{nm} {member} axis 1", N * s + R * s, lambda: t1._fill_axis(member, 1, a))
            run(f"{nm} {member} axis 0", N * s + C * s, lambda: t0._fill_axis(member, 0, a))
        for member in ("ArgMinLastAxis", "ArgMaxLastAxis"):
            run(f"{nm} {member} axis 1", N * s + R * 8, lambda: ti._fill_axis(member, 1, a, True))
        run(f"{nm} FindLastAxis axis 1", N * s + R * 8, lambda: a.Backend.FindLastAxis(7, ti, a))
        del a, b, ta, tb, c, m
    # bool logic
    ta, tb = domain("And", (R, C), dtypes.DN_BOOL), domain("And", (R, C), dtypes.DN_BOOL)
    a, b = wrap(ta), wrap(tb)
    c = wrap(torch.empty((R, C), device="cuda", dtype=torch.bool))
    run("bool Negate", 2 * N, lambda: c.FillNegate(a))
    for member in ("And", "Or", "Xor"):
        run(f"bool {member}", 3 * N, lambda: getattr(c, "Fill" + member)(a, b))
    tb1 = Tensor.empty((R,), dtypes.DN_BOOL, dev)
    tc1 = Tensor.empty((R,), dtypes.DN_I64, dev)
    run("bool AllLastAxis axis 1", N + R, lambda: tb1.FillAllAxis(1, a))
    run("bool AnyLastAxis axis 1", N + R, lambda: tb1.FillAnyAxis(1, a))
    run("bool CountTrueLastAxis axis 1", N + 8 * R, lambda: tc1._fill_axis("CountTrueLastAxis", 1, a, True))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
