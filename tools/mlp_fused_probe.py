import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
from deepnet_b200 import CudaTensor
from mlp_step import init_params, synthetic_batch, train_step
sizes, batch = (784, 4096, 4096, 10), 8192
dev = CudaTensor.dev(); dev.Init(0); dev.SetStream(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(5)
params = [(CudaTensor.ofNumpy(w), CudaTensor.ofNumpy(b)) for w, b in init_params(rng, sizes)]
xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
for _ in range(3):
    train_step(x, t, params, 0.01, fused=True)
torch.cuda.synchronize()
import cProfile, pstats, time
t0 = time.perf_counter()
for _ in range(10):
    train_step(x, t, params, 0.01, fused=True)
host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print("host time per fused step (enqueue only): %.3f ms" % (host * 1e3))
t0 = time.perf_counter()
for _ in range(10):
    train_step(x, t, params, 0.01, fused=False)
host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print("host time per unfused step (enqueue only): %.3f ms" % (host * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    train_step(x, t, params, 0.01, fused=True)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
