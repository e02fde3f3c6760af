#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 Tensor backend (contract: see the task brief / DESIGN.md §5).

Metric (BASELINE.json): element-wise / reduction HBM GB/s (% of B200 peak) at 1/2/4/8 GPUs vs HostTensor.
Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): strided / broadcast element-wise operators on transposed,
broadcast, sliced and reversed views of [16384, 16384] tensors (2^28 elements) in float32, float64 and int32.
One "step" = one pass over that whole case list (33 backend calls).

`value`  = algorithmic bytes of the step, summed over all ranks, / step time (max over ranks), inputs resident in HBM.
`e2e`    = the same case list through the public API from PINNED HOST buffers: every step uploads the inputs
           (dn_memcpy_h2d), runs the 33 calls and downloads the final result of every dtype (the target of the last
           call) and the bool mask; uploads, kernels and downloads overlap on three streams.
Multi-GPU (`--gpus N`, one process per GPU under torchrun): STRONG scaling — the 2^28-element tensors are split along
the leading axis into N equal slabs; a rank holds its slab of every operand AS VIEWED (for
`a.T + b` that is the column block a[:, r0:r1], stored [16384, rows] and used through its transposed view).
Element-wise operators need no collective. A weak-scaling line (one full-size tensor set per rank) is reported beside it.
`sharded_reductions` = BASELINE.json configs[2]: ArgMax + Max over 262144 x 1000 float32 split over the ranks
(strong scaling) through the dn_shard_* entry points — peer-memory stores + flag barrier, no NCCL on the path.

Every number is checked: after the timed region the timed buffers are compared with the HostTensor oracle
(`"verified"`; `--verify` re-runs and checks all 33 calls).

`--impl reference` times the reference's own CPU path for the same case list — the C++ restatement of HostTensor
in oracle/ (the F# original cannot run in this image: no dotnet), with its threading policy, on all host cores
of this box, on a bounded sample ([8192,8192] tensors, 30.5 GB algorithmic per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "elementwise/reduction HBM GB/s (% of B200 peak) at 1/2/4/8 GPUs vs HostTensor"
SIDE = 16384  # 2^28 elements per tensor

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch on the contiguous [16384,16384] case, from the committed
# `ncu --set full` captures (file named per entry); None for kernels without a capture.
NCU_TRAFFIC = {
    "ew_kernel<BinaryF<float,ADD>,VEC=8> (256-bit vector kernel)": (2.147553e9 + 1.041274e9, "profiles/r01b_ew_add_contig256.raw.csv"),
    "ew_kernel<BinaryF<double,ADD>,VEC=4> (256-bit vector kernel)": (4.295071e9 + 2.117457e9, "profiles/r02_ew_add_contig256_f64.raw.csv"),
    "ew_xpose_kernel<BinaryF<float,ADD>> (register transpose)": (2.147507e9 + 1.039083e9, "profiles/r01_ew_xpose_addT.raw.csv"),
    "ew_xpose_kernel<BinaryF<double,ADD>> (register transpose)": (4.295036e9 + 2.110988e9, "profiles/r01_ew_xpose_f64_addT.raw.csv"),
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# The C2 case list, written once against the frontend so that the CUDA arm, the e2e arm, the verification and the
# CPU arm run the same calls. Each entry: (name, algorithmic bytes, callable, kernel tag). Algorithmic bytes: every
# DISTINCT element touched counts once; broadcast operands count their own size.
# ---------------------------------------------------------------------------------------------------------------
def build_cases(T, dt, a, b, c, row, col, mask, has_sin: bool, aT=None, bT=None):
    """a, b, c: [R, C]; row: [1, C]; col: [R, 1]; mask: bool [R, C]; aT, bT: [R, C] transposed views (default: a.T,
    b.T — on a leading-axis shard they are views of the rank's column blocks). Returns the case list of one dtype."""
    from deepnet_b200 import dtypes
    s = dtypes.itemsize(dt)
    R, C = c.Shape
    N = R * C
    nm = dtypes.NAMES[dt]
    aT = a.T if aT is None else aT
    bT = b.T if bT is None else bT
    cs, as_, bs = c[1:, 1:], a[1:, 1:], b[1:, 1:]
    ar0, ar1 = a.reverseAxis(0), a.reverseAxis(1)
    T_ = {"single": "float", "double": "double", "int32": "int"}[nm]
    k_add = f"ew_kernel<BinaryF<{T_},ADD>,VEC={32 // s}> (256-bit vector kernel)"
    k_addT = f"ew_xpose_kernel<BinaryF<{T_},ADD>> (register transpose)"
    cases = [
        (f"{nm} add contiguous", 3 * N * s, lambda: c.FillAdd(a, b), k_add),
        (f"{nm} add a.T + b", 3 * N * s, lambda: c.FillAdd(aT, b), k_addT),
        (f"{nm} add a + row[1,C]", (2 * N + C) * s, lambda: c.FillAdd(a, row), k_add),
        (f"{nm} mul a * col[R,1]", (2 * N + R) * s, lambda: c.FillMultiply(a, col), f"ew_kernel<BinaryF<{T_},MUL>> (vector kernel)"),
        (f"{nm} {'sin' if has_sin else 'abs'}(a.T)", 2 * N * s,
         (lambda: c.FillSin(aT)) if has_sin else (lambda: c.FillAbs(aT)), f"ew_xpose_kernel<UnaryF<{T_}>>"),
        (f"{nm} add a[1:,1:] + b[1:,1:]", 3 * (R - 1) * (C - 1) * s, lambda: cs.FillAdd(as_, bs), k_add),
        (f"{nm} add reverseAxis0(a) + b", 3 * N * s, lambda: c.FillAdd(ar0, b), k_add),
        (f"{nm} add reverseAxis1(a) + b", 3 * N * s, lambda: c.FillAdd(ar1, b), k_add),
        (f"{nm} copy a.T", 2 * N * s, lambda: c.CopyFrom(aT), f"ew_xpose_kernel<CopyF<{8 * s}-bit>>"),
        (f"{nm} less a < b.T -> bool", (2 * s + 1) * N, lambda: mask.FillLess(a, bT), f"ew_xpose_kernel<CompareF<{T_},LESS>>"),
        (f"{nm} ifThenElse(mask, a, b)", (3 * s + 1) * N, lambda: c.FillIfThenElse(mask, a, b), f"ew_kernel<SelectF<{8 * s}-bit>>"),
    ]
    return cases


def dtype_list():
    from deepnet_b200 import dtypes
    # third field: use sin as the unary case. float64 sin is FP64-compute-bound on B200 (reported under
    # other_configs), so the HBM metric uses abs on the transposed view for float64 and int32.
    return [(dtypes.DN_F32, np.float32, True), (dtypes.DN_F64, np.float64, False), (dtypes.DN_I32, np.int32, False)]


def host_inputs(rng, side, npdt, rows=None):
    """a, b: [rows, side]; row: [1, side]; col: [rows, 1] — uniform [-50, 50), ints rounded (Benchmark.fs:105-107)."""
    rows = side if rows is None else rows
    if np.issubdtype(npdt, np.floating):
        mk = lambda shape: rng.uniform(-50, 50, size=shape).astype(npdt)
    else:
        mk = lambda shape: np.rint(rng.uniform(-50, 50, size=shape)).astype(npdt)
    return mk((rows, side)), mk((rows, side)), mk((1, side)), mk((rows, 1))


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons DURING the timed region (NVML; nvidia-smi as a fallback)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []   # (sm_mhz, max_mhz, set(reasons))
        self._stop = threading.Event()
        self._thread = None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self._stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append((float(sm), float(mx), {n for n, b in bits.items() if r & b}))
            self._stop.wait(0.01)

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    timeout=5).decode().strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append((float(f[0]), float(f[1]),
                                     {n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")}))
            except Exception:
                pass
            self._stop.wait(0.1)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)
        sm = [s[0] for s in self.samples]
        reasons = set().union(*[s[2] for s in self.samples]) if self.samples else set()
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(s[1] for s in self.samples) if self.samples else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (HostTensor restatement) on a bounded sample, on ALL host cores whatever launched us
# (torchrun exports OMP_NUM_THREADS=1)
# ---------------------------------------------------------------------------------------------------------------
def use_all_host_cores() -> int:
    from oracle import host_tensor
    lib = host_tensor.api().lib
    lib.dno_set_num_threads.argtypes = [__import__("ctypes").c_int]
    lib.dno_get_num_threads.restype = __import__("ctypes").c_int
    lib.dno_set_num_threads(os.cpu_count() or 1)
    return int(lib.dno_get_num_threads())


def host_of(arr: np.ndarray):
    """HostTensor (oracle device) over `arr` without copying it."""
    from deepnet_b200 import Tensor
    from deepnet_b200 import layout as TL
    from oracle.host_tensor import HostTensor, TensorHostStorage
    return Tensor(TL.newC(arr.shape), TensorHostStorage(arr.reshape(-1), HostTensor.Dev))


def run_cpu_cases(side: int, steps: int, warmup: int):
    from deepnet_b200 import Tensor, dtypes
    from oracle.host_tensor import HostTensor
    threads = use_all_host_cores()
    rng = np.random.default_rng(2)
    per_dtype = []
    total_bytes = 0
    for dt, npdt, has_sin in dtype_list():
        an, bn, rn, cn = host_inputs(rng, side, npdt)
        a, b, row, col = (HostTensor.ofNumpy(x) for x in (an, bn, rn, cn))
        c = Tensor.empty((side, side), dt, HostTensor.Dev)
        mask = Tensor.empty((side, side), dtypes.DN_BOOL, HostTensor.Dev)
        cases = build_cases(Tensor, dt, a, b, c, row, col, mask, has_sin)
        per_dtype.append(cases)
        total_bytes += sum(nb for _, nb, _, _ in cases)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for cases in per_dtype:
            for _, _, fn, _ in cases:
                fn()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return total_bytes / sec / 1e9, sec, total_bytes, threads


# ---------------------------------------------------------------------------------------------------------------
# verification of device buffers against the oracle
# ---------------------------------------------------------------------------------------------------------------
def arrays_equal(h: np.ndarray, c: np.ndarray, rtol: float = 0.0) -> bool:
    """Chunked: bit-exact (NaN == NaN), or rel `rtol` for the one transcendental case."""
    if h.shape != c.shape or h.dtype != c.dtype:
        return False
    hf, cf = h.reshape(-1), c.reshape(-1)
    bits = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[h.dtype.itemsize]
    for lo in range(0, hf.size, 1 << 24):
        hh, cc = hf[lo:lo + (1 << 24)], cf[lo:lo + (1 << 24)]
        if np.array_equal(hh.view(bits), cc.view(bits)):
            continue
        if h.dtype.kind != "f":
            return False
        nan = np.isnan(hh) & np.isnan(cc)
        if rtol == 0.0:
            ok = nan | ((hh == cc) & (np.signbit(hh) == np.signbit(cc)))
        else:
            err = np.abs(hh.astype(np.float64) - cc.astype(np.float64))
            ok = nan | (hh == cc) | (err <= rtol * np.abs(hh.astype(np.float64)) + np.finfo(h.dtype).tiny)
        if not ok.all():
            return False
    return True


def measure_other_configs(dev, torch, peak):
    """C1 / C3 / C4 / C5 of BASELINE.json measured once each on rank 0 (reported beside the headline, not part of it).
    Every entry: median of 5 timings of `reps` back-to-back calls, CUDA events on the launching stream."""
    from deepnet_b200 import CudaTensor, Tensor, dtypes
    stream = torch.cuda.current_stream()
    out = {}

    def w(t, dt):
        return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)

    def timed(fn, reps=4):
        fn()
        ts = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            for _ in range(reps):
                fn()
            e.record(stream)
            e.synchronize()
            ts.append(s.elapsed_time(e) / reps)
        return statistics.median(ts)

    def hbm(name, nbytes, fn, reps=4):
        ms = timed(fn, reps)
        out[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}

    F32, I64, BOOL = dtypes.DN_F32, dtypes.DN_I64, dtypes.DN_BOOL
    # C1: c = a*b + sin(a) over 2^24 elements as three backend calls, then SumLastAxis over 4096x4096
    ta, tb = (torch.rand(4096, 4096, device="cuda") * 100 - 50 for _ in range(2))
    a, b = w(ta, F32), w(tb, F32)
    t1, t2, c = (Tensor.empty((4096, 4096), F32, dev) for _ in range(3))
    o1 = Tensor.empty((4096,), F32, dev)
    n1 = 4096 * 4096 * 4

    def c1():
        t1.FillMultiply(a, b)
        t2.FillSin(a)
        c.FillAdd(t1, t2)
        o1.FillSumAxis(1, c)
    hbm("C1 a*b+sin(a) (3 calls) + SumLastAxis 4096x4096 [small kernels: includes host launch gaps]", 9 * n1 + 4096 * 4, c1, 8)
    from deepnet_b200.fused import trace
    prog_c1 = trace(lambda x, y: x * y + x.sin(), 2)

    def c1_fused():
        c.Backend.FusedElemwise(c, [a, b], prog_c1)
        o1.FillSumAxis(1, c)
    hbm("C1 FUSED a*b+sin(a) (1 call, dn_fused_elemwise) + SumLastAxis 4096x4096 [bytes of the fused form]",
        4 * n1 + 4096 * 4, c1_fused, 8)
    # the same calls captured once into a CUDA graph and replayed: the library's entry points only enqueue work on the
    # caller's stream (no synchronisation, no host allocation), so a host that replays a fixed sequence pays one graph
    # launch instead of one ctypes / P-Invoke call + kernel launch per operator
    try:
        side = torch.cuda.Stream()
        graphs = []
        for fn in (c1, c1_fused):
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                dev.SetStream(side.cuda_stream)
                fn()
            dev.SetStream(stream.cuda_stream)
            graphs.append(g)
        dev.api.call("release_stream", side.cuda_stream)
        hbm("C1 a*b+sin(a) (3 calls) + SumLastAxis 4096x4096, the 4 launches replayed from a CUDA graph", 9 * n1 + 4096 * 4,
            graphs[0].replay, 8)
        hbm("C1 FUSED a*b+sin(a) + SumLastAxis 4096x4096, the 2 launches replayed from a CUDA graph", 4 * n1 + 4096 * 4,
            graphs[1].replay, 8)
        del graphs
    except Exception as exc:  # reported, never fatal for the bench line
        dev.SetStream(stream.cuda_stream)
        out["C1 replayed from a CUDA graph"] = {"error": str(exc)[:200]}
    td = torch.rand(8192, 8192, device="cuda", dtype=torch.float64) * 100 - 50
    dd, dc = w(td, dtypes.DN_F64), Tensor.empty((8192, 8192), dtypes.DN_F64, dev)
    hbm("float64 sin 8192x8192 (FP64-compute-bound on B200, not an HBM kernel)", 2 * 8 * 8192 * 8192, lambda: dc.FillSin(dd))
    del td, dd, dc
    # C3: ArgMaxLastAxis + MaxLastAxis over 262144 x 1000 float32
    tl = torch.rand(262144, 1000, device="cuda") * 100 - 50
    lg = w(tl, F32)
    oi, om = Tensor.empty((262144,), I64, dev), Tensor.empty((262144,), F32, dev)
    nb = 262144 * 1000 * 4
    hbm("C3 ArgMaxLastAxis 262144x1000 f32", nb + 262144 * 8, lambda: oi._fill_axis("ArgMaxLastAxis", 1, lg, True))
    hbm("C3 MaxLastAxis 262144x1000 f32", nb + 262144 * 4, lambda: om.FillMaxAxis(1, lg))
    hbm("C3 SumLastAxis 262144x1000 f32", nb + 262144 * 4, lambda: om.FillSumAxis(1, lg))
    import ctypes as C
    d_ = lambda t: t.Backend._d(t)
    hbm("C3 Max + ArgMax in ONE pass (dn_shard_minmax_arg_last_axis, group = NULL)", nb + 262144 * 12,
        lambda: dev.api.call("shard_minmax_arg_last_axis", None, 0, 1, d_(om), d_(oi), 0, d_(lg)))
    del tl, lg
    # C4: gather / scatter / masked get / trueIdx on 2^26 int64 / bool
    N = 1 << 26
    tsrc = torch.randint(-(1 << 40), 1 << 40, (N,), device="cuda", dtype=torch.int64)
    tidx = torch.randint(0, N, (N,), device="cuda", dtype=torch.int64)
    src, idx = w(tsrc, I64), w(tidx, I64)
    trg = Tensor.empty((N,), I64, dev)
    hbm("C4 Gather 2^26 int64 (random idx)", 3 * 8 * N, lambda: trg.FillGather([idx], src), 2)
    hbm("C4 Scatter 2^26 int64 (random idx)", 5 * 8 * N, lambda: trg.FillScatter([idx], src), 2)
    tmask = torch.rand(N, device="cuda") < 0.5
    mask = w(tmask, BOOL)
    ntrue = int(tmask.sum().item())
    got = Tensor.empty((ntrue,), I64, dev)
    hbm("C4 MaskedGet 2^26 int64 p=0.5", N + 8 * N + 8 * ntrue, lambda: src.Backend.MaskedGet(got, src, [mask]), 2)
    hbm("C4 MaskedSet 2^26 int64 p=0.5", N + 16 * ntrue, lambda: trg.Backend.MaskedSet(trg, [mask], got), 2)
    m2 = mask.reshape((8192, 8192))
    ti = Tensor.empty((ntrue, 2), I64, dev)
    hbm("C4 TrueIndices [8192,8192] p=0.5", N + 16 * ntrue, lambda: ti.Backend.TrueIndices(ti, m2), 2)
    del tsrc, tidx, src, idx, trg, got, ti
    # C5: MLP training step 784-4096-4096-10, batch 8192 (tcgen05 MatMatDot + element-wise + reductions)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from mlp_step import flops_per_step, init_params, synthetic_batch, train_step
    sizes, batch = (784, 4096, 4096, 10), 8192
    rng = np.random.default_rng(5)
    params = [(CudaTensor.ofNumpy(wt), CudaTensor.ofNumpy(bs)) for wt, bs in init_params(rng, sizes)]
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
    fl = flops_per_step(batch, sizes)
    th, tw = torch.randn(8192, 4096, device="cuda"), torch.randn(4096, 4096, device="cuda")
    hh, ww, cc = w(th, F32), w(tw, F32), Tensor.empty((8192, 4096), F32, dev)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16 = float(json.load(f)["bf16_tflops"])
    except Exception:
        bf16 = 1590.0
    # both precisions of float32 MatMatDot (dn_set_math_mode): the fp32-accurate default (3xTF32 on tcgen05) and the
    # opt-in single tf32 pass that north_star's rel 1e-2 tolerance allows
    for mode, label in (("tf32", "DN_MATH_TF32 (opt-in: tf32 inputs, rel 1e-2)"),
                        ("fp32", "DN_MATH_FP32 (default: fp32-accurate, 3xTF32)")):
        dev.SetMathMode(mode)
        ms = timed(lambda: train_step(x, t, params, 1e-3), 3)
        out[f"C5 MLP 784-4096-4096-10 batch 8192 training step, {label}"] = {
            "ms": round(ms, 3), "TFLOP/s": round(fl / ms / 1e9, 1), "gemm_tflop_per_step": round(fl / 1e12, 3)}
        ms_f = timed(lambda: train_step(x, t, params, 1e-3, fused=True), 3)
        out[f"C5 MLP training step with FusedElemwise (same bits, fewer passes over HBM), {label}"] = {
            "ms": round(ms_f, 3), "TFLOP/s": round(fl / ms_f / 1e9, 1)}
        # the GEMM alone (largest layer), against cuBLAS-free denominators: measured bf16 peak / 2 for tf32
        ms = timed(lambda: cc.FillDot(hh, ww.T), 5)
        tf = 2.0 * 8192 * 4096 * 4096 / ms / 1e9
        out[f"C5 MatMatDot 8192x4096x4096 f32, {label}"] = {
            "ms": round(ms, 4), "TFLOP/s": round(tf, 1), "frac_of_tf32_peak": round(tf / (bf16 / 2), 3),
            "tensor_flops_issued_TFLOP/s": round(tf * (3 if mode == "fp32" else 1), 1),
            "tf32_peak_assumed": f"{bf16 / 2:.1f} = measured bf16 peak / 2"}
    dev.SetMathMode("fp32")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--side", type=int, default=SIDE, help="tensor side (default 16384 = 2^28 elements)")
    ap.add_argument("--cpu-side", type=int, default=8192,
                    help="side of the bounded CPU sample ([8192,8192]: ~5 s per step on the GPU box's 16 cores)")
    ap.add_argument("--verify", action="store_true", help="re-run every one of the 33 calls and check it against the oracle")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the timed buffers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the weak-scaling / C1 / C3 / C4 / C5 side measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = (f"C2: element-wise ops on transposed/broadcast/sliced/reversed views of [{args.side},{args.side}] "
                f"(2^{int(np.log2(args.side * args.side))} elements) float32+float64+int32, 33 backend calls per step")

    if args.impl == "reference":
        if rank != 0:
            return 0
        # a step on the [8192,8192] sample is ~5 s on 16 cores: --steps / --warmup are honoured as given
        gbs, sec, nbytes, threads = run_cpu_cases(args.cpu_side, max(1, args.steps), max(0, args.warmup))
        sample = (f"same 33-call case list on [{args.cpu_side},{args.cpu_side}] tensors "
                  f"({nbytes / 1e9:.2f} GB algorithmic per step)")
        line = {
            "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64+i32", "data": "synthetic",
            "config": {"workload": workload, "cpu_sample_side": args.cpu_side,
                       "reference_arm": "HostTensor restatement (C++/OpenMP, oracle/), not .NET; rank 0 only, all host cores"},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from deepnet_b200 import CudaTensor, Tensor, dtypes

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = CudaTensor.dev()
    dev.Init(local_rank)
    stream = torch.cuda.current_stream()
    dev.SetStream(stream.cuda_stream)
    api = dev.api
    side = args.side
    if side % world:
        raise SystemExit(f"--side {side} must be divisible by the number of ranks ({world})")
    rows = side // world                    # STRONG scaling: this rank's slab of the leading axis
    r0 = rank * rows
    torch_dt = {dtypes.DN_F32: torch.float32, dtypes.DN_F64: torch.float64, dtypes.DN_I32: torch.int32}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_ok(ok: bool) -> bool:
        t = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def wrap(t, d):
        return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), d, owner=t)

    # ---- inputs: created on the HOST (pinned), so that the e2e arm uploads and the oracle checks the same data ----
    # per dtype: a, b [rows, side] (this rank's row slab); for N > 1 also the column blocks at, bt [side, rows] that
    # the transposed views read (at N = 1 the transposed views alias a and b themselves)
    rng = np.random.default_rng(1000 + rank)
    sets = []
    for dt, npdt, has_sin in dtype_list():
        an, bn, rown, coln = host_inputs(rng, side, npdt, rows)
        host = {"a": an, "b": bn, "row": rown, "col": coln}
        if world > 1:
            host["at"], host["bt"] = (x.T.copy() for x in host_inputs(rng, side, npdt, rows)[:2])   # [side, rows]
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
        devt = {k: v.to("cuda", non_blocking=True) for k, v in pinned.items()}
        devt["c"] = torch.empty((rows, side), device="cuda", dtype=torch_dt[dt])
        devt["mask"] = torch.empty((rows, side), device="cuda", dtype=torch.bool)
        T_ = {k: wrap(v, dtypes.DN_BOOL if k == "mask" else dt) for k, v in devt.items()}
        aT = T_["at"].T if world > 1 else None
        bT = T_["bt"].T if world > 1 else None
        cases = build_cases(Tensor, dt, T_["a"], T_["b"], T_["c"], T_["row"], T_["col"], T_["mask"], has_sin, aT, bT)
        sets.append({"dt": dt, "npdt": npdt, "has_sin": has_sin, "host": {k: v.numpy() for k, v in pinned.items()},
                     "pinned": pinned, "dev": devt, "T": T_, "cases": cases})
        del host
    torch.cuda.synchronize()
    step_bytes = sum(nb for s in sets for _, nb, _, _ in s["cases"])
    ncalls = sum(len(s["cases"]) for s in sets)

    def step():
        for s in sets:
            for _, _, fn, _ in s["cases"]:
                fn()

    # ---- device-resident arm (the headline `value`) -----------------------------------------------------------
    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dev.LaunchCount()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    launches = dev.LaunchCount() - launches0
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    value = step_bytes * world / (ms_per_step * 1e-3) / 1e9

    # ---- verification: the timed buffers against the HostTensor oracle -----------------------------------------
    # After the timed region c holds ifThenElse(mask, a, b) and mask holds a < b.T (the last two calls of every dtype).
    # --verify additionally re-runs each of the 33 calls once and checks its target.
    verified = None
    if not args.no_verify:
        from oracle.host_tensor import HostTensor
        use_all_host_cores()
        ok = True
        checked = 0
        for s in sets:
            dt, h = s["dt"], s["host"]
            ha, hb, hrow, hcol = (host_of(h[k]) for k in ("a", "b", "row", "col"))
            haT = host_of(h["at"]).T if world > 1 else None
            hbT = host_of(h["bt"]).T if world > 1 else None
            hc = Tensor.empty((rows, side), dt, HostTensor.Dev)
            hmask = Tensor.empty((rows, side), dtypes.DN_BOOL, HostTensor.Dev)
            hcases = build_cases(Tensor, dt, ha, hb, hc, hrow, hcol, hmask, s["has_sin"], haT, hbT)
            if args.verify:
                hc.FillConst(0)
                s["T"]["c"].FillConst(0)
                todo = list(zip(hcases, s["cases"]))
            else:
                todo = list(zip(hcases, s["cases"]))[-2:]
                todo = [(hcase, None) for hcase, _ in todo]      # device results are already in the timed buffers
            for (name, _, hfn, _), dcase in todo:
                hfn()
                if dcase is not None:
                    dcase[2]()
                torch.cuda.synchronize()
                if "-> bool" in name:
                    hh, cc = hmask.Storage.array.reshape(rows, side), s["dev"]["mask"].cpu().numpy()
                else:
                    hh, cc = hc.Storage.array.reshape(rows, side), s["dev"]["c"].cpu().numpy()
                    if "[1:,1:]" in name:
                        hh, cc = np.ascontiguousarray(hh[1:, 1:]), np.ascontiguousarray(cc[1:, 1:])
                ok = ok and arrays_equal(hh, cc, 1e-5 if " sin(" in name else 0.0)
                checked += 1
            if not args.verify:   # the mask is an INPUT of the last call: it must be verified too (done above, [-2])
                pass
        ok = all_ranks_ok(ok)
        verified = {"ok": ok, "calls_checked_per_rank": checked,
                    "what": ("all 33 calls re-run once and compared with the HostTensor oracle" if args.verify else
                             "the timed buffers after the timed region (c = ifThenElse(mask,a,b), mask = a < b.T, "
                             "per dtype) compared with the HostTensor oracle"),
                    "rule": "bit-exact; float32 sin rel 1e-5"}
        if not ok:
            print(json.dumps({"error": "verification against the oracle FAILED", "verified": verified}), file=sys.stderr)

    # ---- per-call timing (CUDA events on the launching stream) for the roofline object; outside the timed region
    per_call = []
    for s in sets:
        for name, nb, fn, ktag in s["cases"]:
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            per_call.append((name, nb, statistics.median(ts), ktag))
    clocks = sampler.stop() if rank == 0 else None
    peak, peak_src = load_peaks()
    # dominant kernel = the kernel TYPE with the largest share of the step (several calls launch the same kernel:
    # contiguous, row-broadcast, reversed and the body of the sliced case all run the 256-bit vector add kernel)
    groups = {}
    for name, nb, ms, ktag in per_call:
        g = groups.setdefault(ktag, {"bytes": 0, "ms": 0.0, "calls": []})
        g["bytes"] += nb
        g["ms"] += ms
        g["calls"].append(name)
    total_call_ms = sum(x[2] for x in per_call)
    dom_tag, dom = max(groups.items(), key=lambda kv: kv[1]["ms"])
    nlaunch = len(dom["calls"])
    traffic, traffic_src = NCU_TRAFFIC.get(dom_tag, (None, None))
    if side != SIDE or world != 1:
        traffic, traffic_src = None, "the committed captures are of the 1-GPU [16384,16384] launch"
    roofline = {
        "bound": "hbm", "kernel": dom_tag, "launches_per_step": nlaunch, "calls": dom["calls"],
        "achieved": dom["bytes"] / dom["ms"] / 1e6, "peak": peak, "unit": "GB/s",
        "frac": dom["bytes"] / dom["ms"] / 1e6 / peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "traffic_note": "DRAM bytes of one launch on the contiguous case, copied from the named ncu capture (not "
                        "measured in this run); algorithmic: 3 x 2^28 x element size. achieved = algorithmic bytes of "
                        "all launches of this kernel in a step / their summed duration (CUDA events, this run)",
        "algorithmic_bytes": dom["bytes"] / nlaunch, "peak_source": peak_src,
        "share_of_step": dom["ms"] / total_call_ms,
        "per_call_gbs": {n: round(nb / ms / 1e6, 1) for n, nb, ms, _ in per_call},
        "per_kernel": {k: {"share_of_step": round(v["ms"] / total_call_ms, 4), "GB/s": round(v["bytes"] / v["ms"] / 1e6, 1),
                           "launches": len(v["calls"])} for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])},
        "step_frac_of_peak": value / world / peak,
    }

    # ---- e2e arm: pinned host buffers -> H2D -> the same 33 calls -> D2H of the final result of every dtype -------
    # Three streams through the C ABI (dn_set_stream is per thread, switched around each group of calls): the uploads
    # of dtype k+1 overlap the kernels of dtype k and the downloads of dtype k-1; ordering by dn_event_record /
    # dn_stream_wait_event. The last call of a dtype (ifThenElse) runs per row block, each block's download starting
    # as soon as that block is done, so only one block's D2H is exposed at the end of the step.
    import ctypes as C
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()

    def mk_event():
        e = C.c_void_p()
        api.call("event_create", C.byref(e))
        return e

    NBLK = 4 if rows % 4 == 0 and rows >= 64 else 1
    blk = rows // NBLK
    h2d = d2h = 0
    for s in sets:
        isz = dtypes.itemsize(s["dt"])
        s["up"] = [(s["T"][k], s["pinned"][k]) for k in ("a", "b", "row", "col", "at", "bt") if k in s["pinned"]]
        s["out_c"] = torch.empty((rows, side), dtype=torch_dt[s["dt"]]).pin_memory()
        s["out_m"] = torch.empty((rows, side), dtype=torch.bool).pin_memory()
        h2d += sum(p.numel() * p.element_size() for _, p in s["up"])
        d2h += rows * side * (isz + 1)
        s["ev_up"], s["ev_free"], s["ev_mask"] = mk_event(), mk_event(), mk_event()
        s["ev_blk"] = [mk_event() for _ in range(NBLK)]
        T_ = s["T"]
        s["last_blocks"] = [(lambda lo=lo, T_=T_: T_["c"][lo:lo + blk].FillIfThenElse(T_["mask"][lo:lo + blk],
                                                                                    T_["a"][lo:lo + blk], T_["b"][lo:lo + blk]))
                            for lo in range(0, rows, blk)]

    def e2e_step():
        for s in sets:
            isz = dtypes.itemsize(s["dt"])
            dev.SetStream(s_h2d.cuda_stream)
            api.call("stream_wait_event", s["ev_free"])      # do not overwrite inputs the previous step still reads
            for t, p in s["up"]:
                api.call("memcpy_h2d", t.Storage.BasePtr(), p.data_ptr(), p.numel() * p.element_size())
            api.call("event_record", s["ev_up"])
            dev.SetStream(stream.cuda_stream)
            api.call("stream_wait_event", s["ev_up"])
            for _, _, fn, _ in s["cases"][:-1]:
                fn()
            api.call("event_record", s["ev_mask"])
            for k, fn in enumerate(s["last_blocks"]):
                fn()
                api.call("event_record", s["ev_blk"][k])
            api.call("event_record", s["ev_free"])
            dev.SetStream(s_d2h.cuda_stream)
            api.call("stream_wait_event", s["ev_mask"])
            api.call("memcpy_d2h_async", s["out_m"].data_ptr(), s["T"]["mask"].Storage.BasePtr(), rows * side)
            for k in range(NBLK):
                api.call("stream_wait_event", s["ev_blk"][k])
                off = k * blk * side * isz
                api.call("memcpy_d2h_async", s["out_c"].data_ptr() + off, s["T"]["c"].Storage.BasePtr() + off, blk * side * isz)
        # the step's results must be on the host before the step counts as finished
        for st in (s_h2d, s_d2h):
            dev.SetStream(st.cuda_stream)
            dev.Synchronize()
        dev.SetStream(stream.cuda_stream)
        dev.Synchronize()

    e2e_steps = max(1, min(args.steps, 3))
    for s in sets:
        api.call("event_record", s["ev_free"])
    e2e_step()
    barrier()
    # several streams are involved, so the step is timed on the host between full synchronisations (every
    # e2e_step ends with dn_sync on all three streams); max over ranks below
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = step_bytes * world / (e2e_ms * 1e-3) / 1e9
    e2e_ok = None
    if not args.no_verify:   # what arrived on the host is what the device-resident arm verified
        e2e_ok = all_ranks_ok(all(arrays_equal(s["out_c"].numpy(), s["dev"]["c"].cpu().numpy()) and
                                  arrays_equal(s["out_m"].numpy(), s["dev"]["mask"].cpu().numpy()) for s in sets))
    for st in (s_h2d, s_d2h):
        api.call("release_stream", st.cuda_stream)

    # ---- free the headline buffers before the side measurements ------------------------------------------------
    for s in sets:
        s.clear()
    del sets
    torch.cuda.empty_cache()

    # ---- weak scaling beside the strong headline (N > 1): one full-size tensor set per rank, 5 steps ---------------
    weak = None
    if world > 1 and not args.no_extra:
        try:
            wsets = []
            g = torch.Generator(device="cuda").manual_seed(17 + rank)
            for dt, npdt, has_sin in dtype_list():
                if dt == dtypes.DN_I32:
                    mk = lambda shape: torch.randint(-50, 50, shape, device="cuda", dtype=torch.int32, generator=g)
                else:
                    mk = lambda shape: torch.rand(shape, device="cuda", dtype=torch_dt[dt], generator=g) * 100 - 50
                ts_ = [mk((side, side)), mk((side, side)), mk((1, side)), mk((side, 1)),
                       torch.empty((side, side), device="cuda", dtype=torch_dt[dt])]
                tm = torch.empty((side, side), device="cuda", dtype=torch.bool)
                a, b, row, col, c = (wrap(t, dt) for t in ts_)
                wsets.append((build_cases(Tensor, dt, a, b, c, row, col, wrap(tm, dtypes.DN_BOOL), has_sin), ts_, tm))
            wbytes = sum(nb for cs, _, _ in wsets for _, nb, _, _ in cs)

            def wstep():
                for cs, _, _ in wsets:
                    for _, _, fn, _ in cs:
                        fn()
            for _ in range(3):
                wstep()
            barrier()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(stream)
            for _ in range(5):
                wstep()
            w1.record(stream)
            barrier()
            wms = max_over_ranks(w0.elapsed_time(w1)) / 5
            weak = {"scaling": "weak", "value": wbytes * world / (wms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": wms,
                    "steps": 5, "per_rank": f"[{side},{side}] x 3 dtypes"}
            del wsets
            torch.cuda.empty_cache()
        except Exception as ex:
            weak = {"error": repr(ex)}

    # ---- sharded reductions (BASELINE.json configs[2]): ArgMaxLastAxis + MaxLastAxis over a
    # 262144 x 1000 float32 tensor split along dim 0 over the ranks (STRONG scaling) through dn_shard_*: every rank's
    # reduction kernel stores its rows into all ranks' results over NVLink and signals; no NCCL on the path (the
    # 64-byte window handles are exchanged once through torch.distributed). Timed on the device, max over ranks.
    sharded = None
    try:
        sharded = measure_sharded_c3(dev, torch, dist, rank, world, local_rank, barrier, max_over_ranks, all_ranks_ok,
                                     not args.no_verify)
    except Exception as ex:
        sharded = {"error": repr(ex)}
    dev.SetStream(stream.cuda_stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32+f64+i32", "data": "synthetic",
        "config": {"workload": workload, "calls_per_step": ncalls, "algorithmic_bytes_per_step_per_gpu": step_bytes,
                   "l2": "inputs (0.1-2 GiB each) are far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"leading-axis shards x{world} of the 2^28-element tensors ({rows} rows per rank), "
                                  "no collective"},
        "frac_of_peak": value / world / peak, "peak_gbs": peak, "peak_source": peak_src,
        "verified": verified,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "arrived_intact": e2e_ok,
                "rule": "per step and rank: every input of the 33 calls is uploaded from pinned host memory, all 33 "
                        "calls run, and the final result of every dtype (target of the last call) plus its bool "
                        "mask are downloaded; the 10 intermediate results per dtype are overwritten on the device "
                        "and never leave it"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "weak_scaling": weak,
        "sharded_reductions": sharded,
    }
    if world == 1 and not args.no_extra:
        try:
            torch.cuda.empty_cache()
            line["other_configs"] = measure_other_configs(dev, torch, peak)
        except Exception as ex:  # the side measurements must never take the headline down
            line["other_configs"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        gbs, sec, nbytes, threads = run_cpu_cases(args.cpu_side, 2, 1)
        line["cpu_baseline"] = {
            "value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"same 33-call case list on [{args.cpu_side},{args.cpu_side}] tensors "
                      f"({nbytes / 1e9:.2f} GB algorithmic per step, {sec:.2f} s per step); HostTensor restatement "
                      f"(C++/OpenMP) with the reference's threading policy, not .NET"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_sharded_c3(dev, torch, dist, rank, world, local_rank, barrier, max_over_ranks, all_ranks_ok, verify):
    import ctypes as C
    from deepnet_b200 import CudaTensor, dtypes
    from deepnet_b200.shard import ShardGroup, slab
    R3, C3 = 262144, 1000
    api = dev.api

    def exchange(blob: bytes):
        if world == 1:
            return [blob]
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    grp = ShardGroup.multi_process(dev, rank, world, local_rank, exchange, heap_bytes=64 << 20)
    st = torch.cuda.Stream()
    dev.SetStream(st.cuda_stream)
    grp.set_stream(rank, st.cuda_stream)
    b3, c3 = slab(R3, rank, world)
    rng = np.random.default_rng(3)                       # every rank draws the same full tensor, keeps its slab
    x = rng.uniform(-50, 50, size=(R3, C3)).astype(np.float32)
    tl = torch.from_numpy(x[b3:b3 + c3]).to("cuda")
    lg = CudaTensor.usingPtr(tl.data_ptr(), (c3, C3), dtypes.DN_F32, owner=tl)
    torch.cuda.synchronize()
    # two sets of full results, used alternately (a target is reused only after another collective in between)
    oi = [grp.alloc(rank, (R3,), dtypes.DN_I64) for _ in range(2)]
    om = [grp.alloc(rank, (R3,), dtypes.DN_F32) for _ in range(2)]
    d = lambda t: t.Backend._d(t)
    g, dl = grp._g, d(lg)
    doi, dom = [d(t) for t in oi], [d(t) for t in om]
    f_arg, f_red, f_fused = api._shard_arg_reduce_last_axis, api._shard_reduce_last_axis, api._shard_minmax_arg_last_axis
    DN_MAX, DN_ARG_MAX = 3, 1

    f_start, f_end = api._shard_group_start, api._shard_group_end

    def two_calls(k):
        # one bracket per step: the two kernels launch back to back and share ONE wait at dn_shard_group_end
        api.check(f_start(g))
        api.check(f_arg(g, rank, DN_ARG_MAX, doi[k & 1], b3, dl))
        api.check(f_red(g, rank, DN_MAX, dom[k & 1], b3, dl))
        api.check(f_end(g))

    def one_call(k):
        api.check(f_fused(g, rank, DN_ARG_MAX, dom[k & 1], doi[k & 1], b3, dl))

    def timed(fn, reps=20):
        for k in range(4):
            fn(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(reps):
            fn(k)
        e1.record(st)
        st.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    ms2 = timed(two_calls)
    ok2 = None
    if verify:
        want_i, want_v = x.argmax(axis=1), x.max(axis=1)
        ok2 = all_ranks_ok(bool((oi[1].toNumpy() == want_i).all() and (om[1].toNumpy() == want_v).all()))
    for t in oi + om:
        t.FillConst(0)
    st.synchronize()
    barrier()
    ms1 = timed(one_call)
    ok1 = None
    if verify:
        ok1 = all_ranks_ok(bool((oi[1].toNumpy() == want_i).all() and (om[1].toNumpy() == want_v).all()))
    nbytes = R3 * C3 * 4
    out = {"workload": f"C3 ArgMaxLastAxis + MaxLastAxis over [{R3},{C3}] float32 split into {world} leading-axis "
                       f"slab(s) (strong scaling); every rank ends with the full [{R3}] results",
           "mechanism": "dn_shard_*: reduction kernel stores its rows into every rank's result (peer memory over "
                        "NVLink / CUDA IPC), last CTA signals, stream waits on the flags; no NCCL launch",
           "two_calls": {"ms": ms2, "GB/s": (2 * nbytes + R3 * 12) / ms2 / 1e6, "verified": ok2,
                         "what": "dn_shard_group_start; dn_shard_arg_reduce_last_axis; dn_shard_reduce_last_axis; "
                                 "dn_shard_group_end (the source is read twice; one wait per step)"},
           "one_pass": {"ms": ms1, "GB/s": (nbytes + R3 * 12) / ms1 / 1e6, "verified": ok1,
                        "what": "dn_shard_minmax_arg_last_axis (Max and ArgMax in one pass over the source)"},
           "nvlink_bytes_per_rank_per_step": (world - 1) * c3 * 12,
           "scaling": "strong"}
    dev.Synchronize()
    barrier()
    grp.close()
    api.call("release_stream", st.cuda_stream)
    return out


if __name__ == "__main__":
    sys.exit(main())
