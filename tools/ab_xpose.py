"""A/B timing of the transposed-view element-wise kernels with two builds of the library on the SAME box (clocks and power caps differ from
box to box by a few percent, which is the size of the effects being chased):
    python tools/ab_reduce.py build/ab/libdeepnet_b200_r01.so deepnet_b200/lib/libdeepnet_b200.so
Each library runs in its own process, alternating, three rounds; CUDA-event timing of 20 back-to-back calls."""
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
R = C = 16384
for dt, tdt, s, nm in ((dtypes.DN_F32, torch.float32, 4, "f32"), (dtypes.DN_I32, torch.int32, 4, "i32"), (dtypes.DN_F64, torch.float64, 8, "f64")):
    if tdt == torch.int32:
        ta = torch.randint(-50, 50, (R, C), device="cuda", dtype=tdt); tb = torch.randint(-50, 50, (R, C), device="cuda", dtype=tdt)
    else:
        ta = torch.rand(R, C, device="cuda", dtype=tdt) * 100 - 50; tb = torch.rand(R, C, device="cuda", dtype=tdt) * 100 - 50
    tc = torch.empty(R, C, device="cuda", dtype=tdt)
    a, b, c = w(ta, dt), w(tb, dt), w(tc, dt)
    N = R * C
    out[f"{nm} a.T + b"] = 3 * N * s / timed(lambda: c.FillAdd(a.T, b), 5) / 1e6
    out[f"{nm} copy a.T"] = 2 * N * s / timed(lambda: c.CopyFrom(a.T), 5) / 1e6
    out[f"{nm} abs(a.T)"] = 2 * N * s / timed(lambda: c.FillAbs(a.T), 5) / 1e6
    tm = torch.empty(R, C, device="cuda", dtype=torch.bool); m = w(tm, dtypes.DN_BOOL)
    out[f"{nm} a < b.T"] = (2 * s + 1) * N / timed(lambda: m.FillLess(a, b.T), 5) / 1e6
    out[f"{nm} a + b (contiguous)"] = 3 * N * s / timed(lambda: c.FillAdd(a, b), 5) / 1e6
    del ta, tb, tc, tm
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
