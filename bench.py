#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 Tensor backend (contract: see the task brief / DESIGN.md §Measurement).

Metric (BASELINE.json): element-wise / reduction HBM GB/s (% of B200 peak) at 1/2/4/8 GPUs vs HostTensor.
Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): strided / broadcast element-wise operators on transposed,
broadcast, sliced and reversed views of [16384, 16384] tensors (2^28 elements) in float32, float64 and int32.
One "step" = one pass over that whole case list (33 backend calls). `value` = algorithmic bytes of the step,
summed over all ranks, divided by the step time (max over ranks) — inputs resident in HBM. `e2e` = the same case
list driven through the public API from PINNED HOST buffers: every step copies the inputs host->device, runs the
calls, and copies one result per dtype back.

Multi-GPU (`--gpus N`, launched under torchrun): the leading axis is sharded, every rank owns one [16384, 16384]
slab of every operand (weak scaling); element-wise operators need no collective.

`--impl reference` times the reference's own CPU path for the same case list — the C++ restatement of HostTensor
in oracle/ (the F# original cannot run in this image: no dotnet), with its threading policy, on the host cores
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

NCU_F64_ADD_TRAFFIC = 4.294980e9 + 2.118896e9  # profiles/r01c_ew_add_contig256_f64.raw.csv
METRIC = "elementwise/reduction HBM GB/s (% of B200 peak) at 1/2/4/8 GPUs vs HostTensor"
SIDE = 16384  # 2^28 elements per tensor


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# The C2 case list, written once against the frontend so that the CUDA arm, the e2e arm and the CPU arm run the
# same calls. Each entry: (name, algorithmic bytes, callable).  Algorithmic bytes: every DISTINCT element touched
# counts once; broadcast operands count their own size.
# ---------------------------------------------------------------------------------------------------------------
def build_cases(T, dt, a, b, c, row, col, mask, has_sin: bool):
    """a, b, c: [R, C]; row: [1, C]; col: [R, 1]; mask: bool [R, C]. Returns the case list for one dtype."""
    from deepnet_b200 import dtypes
    s = dtypes.itemsize(dt)
    R, C = a.Shape
    N = R * C
    nm = dtypes.NAMES[dt]
    cs, as_, bs = c[1:, 1:], a[1:, 1:], b[1:, 1:]
    ar0, ar1, aT = a.reverseAxis(0), a.reverseAxis(1), a.T
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
        (f"{nm} less a < b.T -> bool", (2 * s + 1) * N, lambda: mask.FillLess(a, b.T), f"ew_xpose_kernel<CompareF<{T_},LESS>>"),
        (f"{nm} ifThenElse(mask, a, b)", (3 * s + 1) * N, lambda: c.FillIfThenElse(mask, a, b), f"ew_kernel<SelectF<{8 * s}-bit>>"),
    ]
    return cases


def dtype_list():
    from deepnet_b200 import dtypes
    # third field: use sin as the unary case. float64 sin is FP64-compute-bound on B200 (reported under
    # other_configs), so the HBM metric uses abs on the transposed view for float64 and int32.
    return [(dtypes.DN_F32, np.float32, True), (dtypes.DN_F64, np.float64, False), (dtypes.DN_I32, np.int32, False)]


def host_inputs(rng, side, npdt):
    if np.issubdtype(npdt, np.floating):
        mk = lambda shape: rng.uniform(-50, 50, size=shape).astype(npdt)
    else:
        mk = lambda shape: np.rint(rng.uniform(-50, 50, size=shape)).astype(npdt)
    return mk((side, side)), mk((side, side)), mk((1, side)), mk((side, 1))


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
# CPU arm: the oracle (HostTensor restatement) on a bounded sample
# ---------------------------------------------------------------------------------------------------------------
def run_cpu_cases(side: int, steps: int, warmup: int):
    from deepnet_b200 import Tensor, dtypes
    from oracle.host_tensor import HostTensor
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
    return total_bytes / sec / 1e9, sec, total_bytes


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
    # C5: MLP training step 784-4096-4096-10, batch 8192 (tcgen05 TF32 MatMatDot + element-wise + reductions)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from mlp_step import flops_per_step, init_params, synthetic_batch, train_step
    sizes, batch = (784, 4096, 4096, 10), 8192
    rng = np.random.default_rng(5)
    params = [(CudaTensor.ofNumpy(wt), CudaTensor.ofNumpy(bs)) for wt, bs in init_params(rng, sizes)]
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
    ms = timed(lambda: train_step(x, t, params, 1e-3), 3)
    fl = flops_per_step(batch, sizes)
    out["C5 MLP 784-4096-4096-10 batch 8192 training step"] = {"ms": round(ms, 3), "TFLOP/s": round(fl / ms / 1e9, 1),
                                                              "gemm_tflop_per_step": round(fl / 1e12, 3)}
    ms_f = timed(lambda: train_step(x, t, params, 1e-3, fused=True), 3)
    out["C5 MLP training step with FusedElemwise (same bits, fewer passes over HBM)"] = {
        "ms": round(ms_f, 3), "TFLOP/s": round(fl / ms_f / 1e9, 1)}
    # the GEMM alone (largest layer), against cuBLAS-free denominators: measured bf16 peak / 2 for tf32
    th, tw = torch.randn(8192, 4096, device="cuda"), torch.randn(4096, 4096, device="cuda")
    hh, ww, cc = w(th, F32), w(tw, F32), Tensor.empty((8192, 4096), F32, dev)
    ms = timed(lambda: cc.FillDot(hh, ww.T), 5)
    tf = 2.0 * 8192 * 4096 * 4096 / ms / 1e9
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16 = float(json.load(f)["bf16_tflops"])
    except Exception:
        bf16 = 1590.0
    out["C5 MatMatDot 8192x4096x4096 f32 (tcgen05 kind::tf32)"] = {
        "ms": round(ms, 4), "TFLOP/s": round(tf, 1), "frac_of_tf32_peak": round(tf / (bf16 / 2), 3),
        "tf32_peak_assumed": f"{bf16 / 2:.1f} = measured bf16 peak / 2"}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1/C3/C4 side measurements")
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
        warm = max(1, min(args.warmup, 2))
        steps = max(1, min(args.steps, 5))
        gbs, sec, nbytes = run_cpu_cases(args.cpu_side, steps, warm)
        sample = (f"same 33-call case list on [{args.cpu_side},{args.cpu_side}] tensors "
                  f"({nbytes / 1e9:.2f} GB algorithmic per step)")
        line = {
            "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64+i32", "data": "synthetic",
            "config": {"workload": workload, "reference_arm": "HostTensor restatement (C++/OpenMP, oracle/), not .NET"},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
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
    side = args.side
    torch_dt = {dtypes.DN_F32: torch.float32, dtypes.DN_F64: torch.float64, dtypes.DN_I32: torch.int32}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------------------
    rng = np.random.default_rng(1000 + rank)
    per_dtype, keep = [], []
    host_sets = []
    for dt, npdt, has_sin in dtype_list():
        g = torch.Generator(device="cuda").manual_seed(17 + rank)
        if dt == dtypes.DN_I32:
            mk = lambda shape: torch.randint(-50, 50, shape, device="cuda", dtype=torch.int32, generator=g)
        else:
            mk = lambda shape: torch.rand(shape, device="cuda", dtype=torch_dt[dt], generator=g) * 100 - 50
        ta, tb, trow, tcol = mk((side, side)), mk((side, side)), mk((1, side)), mk((side, 1))
        tc = torch.empty((side, side), device="cuda", dtype=torch_dt[dt])
        tm = torch.empty((side, side), device="cuda", dtype=torch.bool)
        keep += [ta, tb, trow, tcol, tc, tm]
        w = lambda t, d=dt: CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), d, owner=t)
        a, b, row, col, c = w(ta), w(tb), w(trow), w(tcol), w(tc)
        mask = CudaTensor.usingPtr(tm.data_ptr(), (side, side), dtypes.DN_BOOL, owner=tm)
        per_dtype.append((dt, build_cases(Tensor, dt, a, b, c, row, col, mask, has_sin), (a, b, c)))
    step_bytes = sum(nb for _, cases, _ in per_dtype for _, nb, _, _ in cases)
    ncalls = sum(len(cases) for _, cases, _ in per_dtype)

    def step():
        for _, cases, _ in per_dtype:
            for _, _, fn, _ in cases:
                fn()

    for _ in range(max(3, args.warmup)):
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
    total_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = step_bytes * world / (ms_per_step * 1e-3) / 1e9

    # per-call timing (CUDA events on the launching stream) for the roofline object; outside the timed region
    per_call = []
    for _, cases, _ in per_dtype:
        for name, nb, fn, ktag in cases:
            ts = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                fn()
                e.record(stream)
                e.synchronize()
                ts.append(s.elapsed_time(e))
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
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel on the contiguous case, from the
    # `ncu --set full` captures under profiles/ (None for kernels that have not been captured)
    ncu_traffic = {
        "ew_kernel<BinaryF<float,ADD>,VEC=8> (256-bit vector kernel)": 2.147553e9 + 1.041274e9,   # r01b_ew_add_contig256
        "ew_kernel<BinaryF<double,ADD>,VEC=4> (256-bit vector kernel)": NCU_F64_ADD_TRAFFIC,       # r01c_ew_add_contig256_f64
        "ew_xpose_kernel<BinaryF<float,ADD>> (register transpose)": 2.147507e9 + 1.039083e9,      # r01_ew_xpose_addT
        "ew_xpose_kernel<BinaryF<double,ADD>> (register transpose)": 4.295036e9 + 2.110988e9,     # r01_ew_xpose_f64_addT
    }
    roofline = {
        "bound": "hbm", "kernel": dom_tag, "launches_per_step": nlaunch, "calls": dom["calls"],
        "achieved": dom["bytes"] / dom["ms"] / 1e6, "peak": peak, "unit": "GB/s",
        "frac": dom["bytes"] / dom["ms"] / 1e6 / peak,
        "traffic": ncu_traffic.get(dom_tag) if side == SIDE else None,
        "traffic_note": "DRAM bytes of one launch on the contiguous case (algorithmic: 3 x 2^28 x element size); "
                        "achieved = algorithmic bytes of all launches of this kernel in a step / their summed duration",
        "algorithmic_bytes": dom["bytes"] / nlaunch, "peak_source": peak_src,
        "share_of_step": dom["ms"] / total_call_ms,
        "per_call_gbs": {n: round(nb / ms / 1e6, 1) for n, nb, ms, _ in per_call},
        "per_kernel": {k: {"share_of_step": round(v["ms"] / total_call_ms, 4), "GB/s": round(v["bytes"] / v["ms"] / 1e6, 1),
                           "launches": len(v["calls"])} for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])},
        "step_frac_of_peak": (step_bytes / (ms_per_step * 1e-3) / 1e9) / peak if world == 1 else value / world / peak,
    }

    # ---- e2e arm: pinned host buffers -> H2D -> same calls -> D2H of one result per dtype ---------------------
    # Pinned host buffers (8 GiB per rank at the default size): float32 and int32 share one pair of input buffers
    # (the int32 cases upload the same bytes), float64 has its own pair, one download buffer is shared.
    e2e_side = side
    pinned = []
    h2d = d2h = 0
    out_buf = torch.empty(e2e_side * e2e_side * 8, dtype=torch.uint8).pin_memory()
    shared32 = None
    for dt, npdt, has_sin in dtype_list():
        isz = dtypes.itemsize(dt)
        if isz == 4 and shared32 is not None:
            hp = [x.view(torch_dt[dt]) for x in shared32]
        else:
            an, bn, _, _ = host_inputs(rng, e2e_side, npdt)
            hp = [torch.from_numpy(x).pin_memory() for x in (an, bn)]
            del an, bn
            if isz == 4:
                shared32 = hp
        out = out_buf[: e2e_side * e2e_side * isz].view(torch_dt[dt]).view(e2e_side, e2e_side)
        pinned.append((hp, out))
        h2d += sum(x.numel() * x.element_size() for x in hp)
        d2h += out.numel() * out.element_size()
    api = dev.api
    # Three streams through the C ABI (dn_set_stream is per thread, switched around each group of calls): uploads
    # of dtype k+1 overlap the kernels of dtype k and the download of dtype k-1. Ordering is by events
    # (dn_event_record / dn_stream_wait_event); the step ends with dn_sync on every stream.
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    import ctypes as C

    def mk_event():
        e = C.c_void_p()
        api.call("event_create", C.byref(e))
        return e

    ev_up = [mk_event() for _ in pinned]
    ev_done = [mk_event() for _ in pinned]
    ev_free = [mk_event() for _ in pinned]   # previous step's kernels done with a, b of this dtype

    def e2e_step():
        for k, ((dt, cases, (a, b, c)), (hp, out)) in enumerate(zip(per_dtype, pinned)):
            isz = dtypes.itemsize(dt)
            dev.SetStream(s_h2d.cuda_stream)
            api.call("stream_wait_event", ev_free[k])      # do not overwrite inputs still being read
            api.call("memcpy_h2d", a.Storage.BasePtr(), hp[0].data_ptr(), hp[0].numel() * isz)
            api.call("memcpy_h2d", b.Storage.BasePtr(), hp[1].data_ptr(), hp[1].numel() * isz)
            api.call("event_record", ev_up[k])
            dev.SetStream(stream.cuda_stream)
            api.call("stream_wait_event", ev_up[k])
            for _, _, fn, _ in cases:
                fn()
            api.call("event_record", ev_done[k])
            api.call("event_record", ev_free[k])
            dev.SetStream(s_d2h.cuda_stream)
            api.call("stream_wait_event", ev_done[k])
            api.call("memcpy_d2h_async", out.data_ptr(), c.Storage.BasePtr(), out.numel() * isz)
        # the step's results must be on the host before the step counts as finished
        for st in (s_h2d, s_d2h):
            dev.SetStream(st.cuda_stream)
            dev.Synchronize()
        dev.SetStream(stream.cuda_stream)
        dev.Synchronize()

    e2e_steps = max(1, min(args.steps, 3))
    for k in range(len(pinned)):
        api.call("event_record", ev_free[k])
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
    t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = step_bytes * world / (e2e_ms * 1e-3) / 1e9

    # ---- sharded reductions (BASELINE.json configs[2]): ArgMaxLastAxis + MaxLastAxis over a
    # 262144 x 1000 float32 tensor split along dim 0 over the ranks (strong scaling), outputs combined by an NCCL
    # all-gather; timed on the device, max over ranks. Reported beside the headline.
    sharded = None
    try:
        from deepnet_b200.shard import LeadingAxisSharding, slab
        TD = {torch.float32: dtypes.DN_F32, torch.int64: dtypes.DN_I64, torch.bool: dtypes.DN_BOOL}
        sh = LeadingAxisSharding(lambda t: CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), TD[t.dtype], owner=t),
                                 torch.device("cuda", local_rank))
        C3 = 1000

        def c3_time(R3):
            b3, c3 = slab(R3, rank, world)
            tl = torch.rand(c3, C3, device="cuda") * 100 - 50
            lg = CudaTensor.usingPtr(tl.data_ptr(), (c3, C3), dtypes.DN_F32, owner=tl)

            def c3_step():
                sh.reduce_axis("ArgMaxLastAxis", lg, 1, R3)
                sh.reduce_axis("MaxLastAxis", lg, 1, R3)
            for _ in range(3):
                c3_step()
            barrier()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps3 = 10
            s3.record(stream)
            for _ in range(reps3):
                c3_step()
            e3.record(stream)
            barrier()
            t3 = torch.tensor([s3.elapsed_time(e3) / reps3], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            ms3 = float(t3.item())
            return ms3, (2 * R3 * C3 * 4 + R3 * 12) / ms3 / 1e6

        ms_s, gbs_s = c3_time(262144)            # strong: the config's tensor split over the ranks
        ms_w, gbs_w = c3_time(262144 * world)    # weak: one config-sized slab per rank
        sharded = {"workload": f"C3 ArgMaxLastAxis + MaxLastAxis over [R,{C3}] float32, {world} leading-axis shard(s), "
                               "outputs all-gathered in place (NCCL)",
                   "strong": {"rows": 262144, "ms": ms_s, "GB/s": gbs_s},
                   "weak": {"rows": 262144 * world, "ms": ms_w, "GB/s": gbs_w}}
    except Exception as ex:
        sharded = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64+i32", "data": "synthetic",
        "config": {"workload": workload, "calls_per_step": ncalls, "algorithmic_bytes_per_step_per_gpu": step_bytes,
                   "l2": "inputs (1-2 GiB each) are far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"leading-axis shards x{world}, no collective"},
        "frac_of_peak": value / world / peak, "peak_gbs": peak, "peak_source": peak_src,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "sharded_reductions": sharded,
    }
    if world == 1 and not args.no_extra:
        try:
            del per_dtype, keep, pinned
            torch.cuda.empty_cache()
            line["other_configs"] = measure_other_configs(dev, torch, peak)
        except Exception as ex:  # the side measurements must never take the headline down
            line["other_configs"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        gbs, sec, nbytes = run_cpu_cases(args.cpu_side, 2, 1)
        line["cpu_baseline"] = {
            "value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"same 33-call case list on [{args.cpu_side},{args.cpu_side}] tensors "
                      f"({nbytes / 1e9:.2f} GB algorithmic per step, {sec:.2f} s per step); HostTensor restatement "
                      f"(C++/OpenMP) with the reference's threading policy, not .NET"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
