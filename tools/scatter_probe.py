"""Probe: scatter throughput as a function of the target size (L2-resident vs DRAM-resident atomics), the binned
path at full size, and its per-kernel times."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepnet_b200 import CudaTensor, Tensor, dtypes
dev = CudaTensor.dev(); dev.Init(0)
stream = torch.cuda.current_stream(); dev.SetStream(stream.cuda_stream)
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
src_t = torch.randint(-(1 << 40), 1 << 40, (N,), device="cuda", dtype=torch.int64)
src = w(src_t, dtypes.DN_I64)
for logt in (20, 21, 22, 23, 24, 26):
    T = 1 << logt
    idx_t = torch.randint(0, T, (N,), device="cuda", dtype=torch.int64)
    idx = w(idx_t, dtypes.DN_I64)
    trg = Tensor.empty((T,), dtypes.DN_I64, dev)
    ms = timed(lambda: trg.FillScatter([idx], src))
    ref = torch.zeros(T, device="cuda", dtype=torch.int64).scatter_add_(0, idx_t, src_t)
    ok = bool((torch.from_numpy(trg.toNumpy()).cuda() == ref).all())
    print(f"scatter 2^26 int64 into 2^{logt}: {ms:.3f} ms  exact={ok}", flush=True)
for name, mk in (("permutation", lambda: torch.randperm(N, device="cuda")),
                 ("hot-spot 4096 cells", lambda: torch.randint(0, 4096, (N,), device="cuda") * 16001 % N)):
    idx_t = mk().to(torch.int64)
    idx = w(idx_t, dtypes.DN_I64)
    trg = Tensor.empty((N,), dtypes.DN_I64, dev)
    ms = timed(lambda: trg.FillScatter([idx], src))
    print(f"scatter 2^26 int64 {name}: {ms:.3f} ms", flush=True)
f_t = torch.rand(N, device="cuda"); fsrc = w(f_t, dtypes.DN_F32)
idx_t = torch.randint(0, N, (N,), device="cuda", dtype=torch.int64); idx = w(idx_t, dtypes.DN_I64)
trg = Tensor.empty((N,), dtypes.DN_F32, dev)
print(f"scatter 2^26 float32 random: {timed(lambda: trg.FillScatter([idx], fsrc)):.3f} ms", flush=True)
