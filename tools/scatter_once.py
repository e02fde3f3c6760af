"""One binned Scatter of 2^26 random int64 adds into 2^26 and into 2^20 cells — run under
`ncu --metrics gpu__time_duration.sum` to get the time of each kernel of the pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepnet_b200 import CudaTensor, Tensor, dtypes
dev = CudaTensor.dev(); dev.Init(0)
dev.SetStream(torch.cuda.current_stream().cuda_stream)
def w(t, dt): return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)
N = 1 << 26
src_t = torch.randint(-(1 << 40), 1 << 40, (N,), device="cuda", dtype=torch.int64)
src = w(src_t, dtypes.DN_I64)
for logt in (26, 20):
    T = 1 << logt
    idx_t = torch.randint(0, T, (N,), device="cuda", dtype=torch.int64)
    idx = w(idx_t, dtypes.DN_I64)
    trg = Tensor.empty((T,), dtypes.DN_I64, dev)
    for _ in range(2):
        trg.FillScatter([idx], src)
    torch.cuda.synchronize()
