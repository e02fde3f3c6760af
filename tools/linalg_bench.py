import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
from deepnet_b200 import CudaTensor, Tensor, dtypes
from perf_sweep import wrap, timeit
dev = CudaTensor.dev(); dev.Init(0); dev.SetStream(torch.cuda.current_stream().cuda_stream)
for (B, M, N, K) in [(4096, 32, 32, 32), (1024, 64, 64, 64), (256, 128, 128, 128), (64, 256, 256, 256), (16, 512, 512, 512)]:
    for tdt, dt in ((torch.float32, dtypes.DN_F32), (torch.float64, dtypes.DN_F64)):
        ta, tb = torch.randn(B, M, K, device="cuda", dtype=tdt), torch.randn(B, K, N, device="cuda", dtype=tdt)
        a, b = wrap(ta), wrap(tb)
        c = Tensor.empty((B, M, N), dt, dev)
        ms = timeit(lambda: c.FillDot(a, b), 5)
        tms = timeit(lambda: torch.bmm(ta, tb), 5)
        fl = 2.0 * B * M * N * K
        print(f"batched {dtypes.NAMES[dt]:6s} {B}x[{M},{K}].[{K},{N}]  ours {ms:8.3f} ms {fl/ms/1e9:8.2f} TFLOP/s   torch.bmm {tms:8.3f} ms {fl/tms/1e9:8.2f} TFLOP/s")
# invert
for (B, n) in [(4096, 8), (1024, 32), (256, 64), (64, 128), (16, 200), (4, 512)]:
    for tdt, dt in ((torch.float32, dtypes.DN_F32), (torch.float64, dtypes.DN_F64)):
        tm = torch.randn(B, n, n, device="cuda", dtype=tdt) + torch.eye(n, device="cuda", dtype=tdt) * n
        m = wrap(tm)
        out = Tensor.empty((B, n, n), dt, dev)
        ms = timeit(lambda: out.FillInvert(m), 3)
        tms = timeit(lambda: torch.linalg.inv(tm), 3)
        print(f"invert  {dtypes.NAMES[dt]:6s} {B}x[{n},{n}]  ours {ms:8.3f} ms   torch.linalg.inv {tms:8.3f} ms")
