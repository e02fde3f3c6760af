"""Two-rank NCCL run of the leading-axis sharding on real GPUs (skipped unless >= 2 devices are visible; run with
`gpurun --gpus 2`). Each rank holds a slab on its own B200, reduces it with libdeepnet_b200.so and the partials are
combined through torch.distributed (NCCL) + dn_arg_reduce_combine; results are compared with a single-GPU run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["DN_ROOT"])
import numpy as np, torch, torch.distributed as dist
from deepnet_b200 import CudaTensor, Tensor, dtypes, NotFound
from deepnet_b200.shard import LeadingAxisSharding, slab
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = CudaTensor.dev(); dev.Init(lr); dev.SetStream(torch.cuda.current_stream().cuda_stream)
TD = {torch.float32: dtypes.DN_F32, torch.int64: dtypes.DN_I64, torch.bool: dtypes.DN_BOOL, torch.int32: dtypes.DN_I32,
      torch.float64: dtypes.DN_F64}
wrap = lambda t: CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), TD[t.dtype], owner=t)
sh = LeadingAxisSharding(wrap, torch.device("cuda", lr))
rng = np.random.default_rng(3)
R, C = 4099, 1000
f = rng.uniform(-50, 50, size=(R, C)).astype(np.float32)
f[5, 1] = np.nan; f[4000, 2] = np.nan; f[:, 3] = -np.inf; f[100, 4] = 99.0; f[3000, 4] = 99.0
full = CudaTensor.ofNumpy(f)
b, c = slab(R, rank, world)
loc = CudaTensor.ofNumpy(f[b:b + c])
bad = []
def same(name, got, want):
    g, w = got.toNumpy(), want.toNumpy()
    if not (g.shape == w.shape and ((g == w) | (np.isnan(g.astype(np.float64)) & np.isnan(w.astype(np.float64)))).all()):
        bad.append(name)
for member, fn in [("MaxLastAxis", "maxAxis"), ("MinLastAxis", "minAxis"), ("ArgMaxLastAxis", "argMaxAxis"),
                   ("ArgMinLastAxis", "argMinAxis")]:
    for axis in (0, 1):
        same(f"{member} axis {axis}", sh.reduce_axis(member, loc, axis, R), getattr(full, fn)(axis))
same("find", sh.reduce_axis("FindLastAxis", loc, 0, R, value=99.0), full.findAxis(99.0, 0))
same("whole argmax", sh.reduce_axis("ArgMaxLastAxis", loc.flatten(), 0, R * C), full.flatten().argMaxAxis(0))
got = sh.reduce_axis("SumLastAxis", loc[:, 10:], 1, R).toNumpy(); want = full[:, 10:].sumAxis(1).toNumpy()
if not np.allclose(got, want, rtol=1e-3, atol=1e-1): bad.append("sum axis 1")
mk = rng.uniform(0, 1, size=(R, C)) < 0.3
fm, lm = CudaTensor.ofNumpy(mk), CudaTensor.ofNumpy(mk[b:b + c])
same("trueIdx", sh.true_indices(lm, R), fm.trueIdx())
same("maskedGet", sh.masked_get(loc, lm), full.M(fm))
dist.barrier()
sys.stdout.write(f"RANK{rank}_" + ("OK" if not bad else "FAIL " + ";".join(bad)) + "\n"); sys.stdout.flush()
dist.destroy_process_group()
'''


def test_sharded_reductions_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DN_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert "RANK0_OK" in out.stdout and "RANK1_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
