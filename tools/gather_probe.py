"""Probe: does the L2 fetch granularity (cudaLimitMaxL2FetchGranularity: 32 / 64 / 128 bytes) change the random-access
kernels (Gather, Scatter at 2^26 int64) and does it cost the streaming ones anything? ncu showed ~110 bytes of DRAM
reads per random 8-byte access with the default setting (profiles/r01_gather.raw.csv)."""
import ctypes as C, glob, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepnet_b200 import CudaTensor, Tensor, dtypes
dev = CudaTensor.dev(); dev.Init(0)
stream = torch.cuda.current_stream(); dev.SetStream(stream.cuda_stream)
rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.*"))[0])
LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity
def get():
    v = C.c_size_t(); rc = rt.cudaDeviceGetLimit(C.byref(v), LIMIT); return (rc, v.value)
def w(t, dt): return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)
def timed(fn, reps=3):
    fn(); ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(reps): fn()
        e.record(stream); e.synchronize(); ts.append(s.elapsed_time(e) / reps)
    return statistics.median(ts)
N = 1 << 26
src_t = torch.randint(-(1 << 40), 1 << 40, (N,), device="cuda", dtype=torch.int64); src = w(src_t, dtypes.DN_I64)
idx_t = torch.randint(0, N, (N,), device="cuda", dtype=torch.int64); idx = w(idx_t, dtypes.DN_I64)
perm_t = torch.randperm(N, device="cuda").to(torch.int64); perm = w(perm_t, dtypes.DN_I64)
trg = Tensor.empty((N,), dtypes.DN_I64, dev)
a_t = torch.rand(1 << 27, device="cuda", dtype=torch.float64); b_t = torch.rand(1 << 27, device="cuda", dtype=torch.float64)
a, b = w(a_t, dtypes.DN_F64), w(b_t, dtypes.DN_F64); c = Tensor.empty((1 << 27,), dtypes.DN_F64, dev)
m_t = torch.rand(N, device="cuda") < 0.5; m = w(m_t, dtypes.DN_BOOL)
nt = int(m_t.sum()); got = Tensor.empty((nt,), dtypes.DN_I64, dev)
print("default limit (rc, bytes):", get(), flush=True)
for g in (128, 64, 32, 128):
    rc = rt.cudaDeviceSetLimit(LIMIT, C.c_size_t(g))
    print(f"--- set {g}: rc={rc}, now {get()}", flush=True)
    print(f"gather random      {timed(lambda: trg.FillGather([idx], src)):.3f} ms", flush=True)
    print(f"gather permutation {timed(lambda: trg.FillGather([perm], src)):.3f} ms", flush=True)
    print(f"scatter staged     {timed(lambda: trg.FillScatter([idx], src)):.3f} ms", flush=True)
    print(f"add f64 2^27       {timed(lambda: c.FillAdd(a, b)):.3f} ms", flush=True)
    print(f"maskedGet p=0.5    {timed(lambda: src.Backend.MaskedGet(got, src, [m])):.3f} ms", flush=True)
    print(f"maskedSet p=0.5    {timed(lambda: trg.Backend.MaskedSet(trg, [m], got)):.3f} ms", flush=True)
want = src_t[idx_t]
trg.FillGather([idx], src)
print("gather exact:", bool((torch.from_numpy(trg.toNumpy()).cuda() == want).all()))
