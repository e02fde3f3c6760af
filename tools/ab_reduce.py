"""A/B timing of the C3 reductions with two builds of the library on the SAME box (clocks and power caps differ from
box to box by a few percent, which is the size of the effects being chased):
    python tools/ab_reduce.py build/ab/libdeepnet_b200_r01.so deepnet_b200/lib/libdeepnet_b200.so
Each library runs in its own process, alternating, three rounds; CUDA-event timing of 20 back-to-back calls.
An older build for the left column: `git archive <rev> | tar -x -C build/ab_src && make -C build/ab_src -j &&
cp build/ab_src/deepnet_b200/lib/libdeepnet_b200.so build/ab/libdeepnet_b200_<rev>.so` (build/ is git-ignored but travels
to the GPU box). tools/ab_xpose.py is the same harness for the transposed-view element-wise kernels."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, statistics
sys.path.insert(0, os.environ["DN_ROOT"])
from deepnet_b200 import native
native.product_library_path = lambda: os.environ["DN_AB_LIB"]
import ctypes
_lib = ctypes.CDLL(os.environ["DN_AB_LIB"])
for table in (native._OPERATOR_SIGNATURES, native._DEVICE_SIGNATURES):   # an older build exports fewer entry points
    for name in [n for n in table if not hasattr(_lib, "dn_" + n)]:
        del table[name]
import torch
from deepnet_b200 import CudaTensor, Tensor, dtypes
dev = CudaTensor.dev(); dev.Init(0)
stream = torch.cuda.current_stream(); dev.SetStream(stream.cuda_stream)
def w(t, dt): return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)
def timed(fn, reps=20):
    fn(); ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(reps): fn()
        e.record(stream); e.synchronize(); ts.append(s.elapsed_time(e) / reps)
    return statistics.median(ts)
out = {}
tl = torch.rand(262144, 1000, device="cuda") * 100 - 50
lg = w(tl, dtypes.DN_F32)
oi, om = Tensor.empty((262144,), dtypes.DN_I64, dev), Tensor.empty((262144,), dtypes.DN_F32, dev)
nb = 262144 * 1000 * 4
out["C3 argmax"] = nb / timed(lambda: oi._fill_axis("ArgMaxLastAxis", 1, lg, True)) / 1e6
out["C3 max"] = nb / timed(lambda: om.FillMaxAxis(1, lg)) / 1e6
out["C3 sum"] = nb / timed(lambda: om.FillSumAxis(1, lg)) / 1e6
if hasattr(_lib, "dn_shard_minmax_arg_last_axis"):
    _d = lambda t: t.Backend._d(t)
    out["C3 max+argmax one pass"] = nb / timed(lambda: dev.api.call("shard_minmax_arg_last_axis", None, 0, 1, _d(om), _d(oi), 0, _d(lg))) / 1e6
td = torch.rand(131072, 1000, device="cuda", dtype=torch.float64) * 100 - 50
ld = w(td, dtypes.DN_F64); oid = Tensor.empty((131072,), dtypes.DN_I64, dev)
out["f64 131072x1000 argmax"] = 131072 * 8000 / timed(lambda: oid._fill_axis("ArgMaxLastAxis", 1, ld, True)) / 1e6
ti8 = torch.randint(-100, 100, (262144, 1000), device="cuda", dtype=torch.int8)
l8 = w(ti8, dtypes.DN_I8)
out["i8 262144x1000 argmax"] = 262144 * 1000 / timed(lambda: oi._fill_axis("ArgMaxLastAxis", 1, l8, True)) / 1e6
o8 = Tensor.empty((262144,), dtypes.DN_I8, dev)
out["i8 262144x1000 sum"] = 262144 * 1000 / timed(lambda: o8.FillSumAxis(1, l8)) / 1e6
out["i8 262144x1000 max"] = 262144 * 1000 / timed(lambda: o8.FillMaxAxis(1, l8)) / 1e6
ti16 = torch.randint(-1000, 1000, (262144, 1000), device="cuda", dtype=torch.int16)
l16 = w(ti16, dtypes.DN_I16); o16 = Tensor.empty((262144,), dtypes.DN_I16, dev)
out["i16 262144x1000 max"] = 262144 * 2000 / timed(lambda: o16.FillMaxAxis(1, l16)) / 1e6
out["i16 262144x1000 argmax"] = 262144 * 2000 / timed(lambda: oi._fill_axis("ArgMaxLastAxis", 1, l16, True)) / 1e6
tb = torch.rand(262144, 1000, device="cuda") < 0.5
lb = w(tb, dtypes.DN_BOOL); ob = Tensor.empty((262144,), dtypes.DN_BOOL, dev)
out["bool 262144x1000 countTrue"] = 262144 * 1000 / timed(lambda: oi._fill_axis("CountTrueLastAxis", 1, lb, True)) / 1e6
out["bool 262144x1000 all"] = 262144 * 1000 / timed(lambda: ob.FillAllAxis(1, lb)) / 1e6
tb2 = torch.rand(16384, 16384, device="cuda") < 0.5
lb2 = w(tb2, dtypes.DN_BOOL); oi2 = Tensor.empty((16384,), dtypes.DN_I64, dev)
out["bool 16384^2 countTrue1"] = 16384 * 16384 / timed(lambda: oi2._fill_axis("CountTrueLastAxis", 1, lb2, True)) / 1e6
ti82 = torch.randint(-100, 100, (16384, 16384), device="cuda", dtype=torch.int8)
l82 = w(ti82, dtypes.DN_I8); o82 = Tensor.empty((16384,), dtypes.DN_I8, dev)
out["i8 16384^2 sum1"] = 16384 * 16384 / timed(lambda: o82.FillSumAxis(1, l82)) / 1e6
del tb, tb2, ti16, ti8, ti82
t2 = torch.rand(16384, 16384, device="cuda") * 100 - 50
a2 = w(t2, dtypes.DN_F32); o2 = Tensor.empty((16384,), dtypes.DN_F32, dev); o2i = Tensor.empty((16384,), dtypes.DN_I64, dev)
nb2 = 16384 * 16384 * 4
out["16384^2 f32 sum1"] = nb2 / timed(lambda: o2.FillSumAxis(1, a2)) / 1e6
out["16384^2 f32 max1"] = nb2 / timed(lambda: o2.FillMaxAxis(1, a2)) / 1e6
out["16384^2 f32 argmax1"] = nb2 / timed(lambda: o2i._fill_axis("ArgMaxLastAxis", 1, a2, True)) / 1e6
out["16384^2 f32 sum0"] = nb2 / timed(lambda: o2.FillSumAxis(0, a2)) / 1e6
print(json.dumps(out))
'''

if __name__ == "__main__":
    libs = [os.path.abspath(p) for p in sys.argv[1:]] or [os.path.join(ROOT, "build/ab/libdeepnet_b200_r01.so"),
                                                          os.path.join(ROOT, "deepnet_b200/lib/libdeepnet_b200.so")]
    res = {p: [] for p in libs}
    for _ in range(3):
        for p in libs:
            env = dict(os.environ, DN_ROOT=ROOT, DN_AB_LIB=p)
            out = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True, timeout=300)
            if out.returncode != 0:
                print(p, "FAILED", out.stderr[-1500:])
                continue
            res[p].append(json.loads(out.stdout.strip().splitlines()[-1]))
    keys = []
    for p in libs:
        for k in (res[p][0] if res[p] else {}):
            if k not in keys:
                keys.append(k)
    print(f"{'GB/s (median of rounds)':28s}" + "".join(f"{os.path.basename(p)[:26]:>28s}" for p in libs))
    for k in keys:
        med = lambda p: sorted(r[k] for r in res[p])[len(res[p]) // 2] if res[p] and k in res[p][0] else float("nan")
        print(f"{k:28s}" + "".join(f"{med(p):28.1f}" for p in libs))
