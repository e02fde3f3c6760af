"""Leading-axis sharding through the dn_shard_* entry points of libdeepnet_b200.so (peer-memory stores + flag barrier
inside the producing kernel; include/dn_tensor.h "Multi-GPU"), compared with the UNSHARDED HostTensor oracle
(tests/shard_cases.py).

Both process models are covered, and both also run on a ONE-GPU box, because ranks may share a device:
  * single process driving 2 / 3 / 8 ranks (the F# host's model) — on distinct devices when the box has them;
  * one process per rank (the torchrun model): the window handles travel through files, the windows are mapped
    with CUDA IPC. No torch.distributed, no NCCL anywhere on the path."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240, method="thread")]


def _device_count(cuda_dev) -> int:
    import ctypes as C
    n = C.c_int32()
    cuda_dev.api.call("device_count", C.byref(n))
    return n.value


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_ops_single_process(cuda_dev, world):
    import shard_cases
    from deepnet_b200.shard import ShardGroup
    ndev = _device_count(cuda_dev)
    if world > 4 * ndev:
        # every rank has a stream that may sit in a wait for a peer's signal; streams beyond the device's hardware
        # queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) share a queue, and a kernel queued behind a waiting
        # stream's wait would never send its signal
        pytest.skip("more than 4 ranks per device")
    devices = [r % ndev for r in range(world)]
    grp = ShardGroup.single_process(cuda_dev, devices)
    try:
        def set_device(r):
            cuda_dev.api.call("set_device", devices[r])
        bad = shard_cases.run(grp, set_device)
    finally:
        cuda_dev.api.call("set_device", 0)
        grp.close()
    assert not bad, bad


WORKER = r'''
import os, sys, time
sys.path.insert(0, os.environ["DN_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DN_ROOT"], "tests"))
import ctypes as C
from deepnet_b200 import CudaTensor
from deepnet_b200.shard import ShardGroup
import shard_cases
rank, world, xdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dev = CudaTensor.dev()
n = C.c_int32(); dev.api.call("device_count", C.byref(n))
device = rank % n.value
dev.Init(device)

def exchange(blob):
    with open(os.path.join(xdir, f"h{rank}.tmp"), "wb") as fh:
        fh.write(blob)
    os.rename(os.path.join(xdir, f"h{rank}.tmp"), os.path.join(xdir, f"h{rank}"))
    out = []
    for r in range(world):
        p = os.path.join(xdir, f"h{r}")
        t0 = time.time()
        while not os.path.exists(p):
            if time.time() - t0 > 120: raise RuntimeError("handle exchange timed out")
            time.sleep(0.01)
        out.append(open(p, "rb").read())
    return out

grp = ShardGroup.multi_process(dev, rank, world, device, exchange)
bad = shard_cases.run(grp, lambda r: None)
sys.stdout.write(f"RANK{rank}_" + ("OK" if not bad else "FAIL " + ";".join(bad)) + "\n"); sys.stdout.flush()
# keep the window mapped until every rank has finished with it
open(os.path.join(xdir, f"done{rank}"), "w").close()
t0 = time.time()
while not all(os.path.exists(os.path.join(xdir, f"done{r}")) for r in range(world)) and time.time() - t0 < 120:
    time.sleep(0.01)
grp.close()
'''


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ops_one_process_per_rank(cuda_dev, tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DN_ROOT=ROOT)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(tmp_path)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (out, err) in enumerate(outs):
        assert f"RANK{r}_OK" in out, (out[-2000:], err[-3000:])
