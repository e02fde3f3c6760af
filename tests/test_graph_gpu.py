"""The entry points only enqueue work on the caller's stream — no synchronisation, no blocking allocation — so a fixed
sequence of backend calls can be captured into a CUDA graph and replayed (the B200-side answer to the reference's
per-call overhead: reflection-built argument arrays and one cuLaunchKernel per operator, CudaKernels.fs:180-262).
Captured here: the C1 chain (Multiply, Sin, Add, SumLastAxis), a strided / broadcast chain, ArgMax + Max and the fused
one-pass form; replays run on NEW input contents and are compared with the HostTensor oracle."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes
from helpers import HostTensor

pytestmark = pytest.mark.gpu


def test_backend_calls_replay_from_a_cuda_graph(cuda_dev):
    import torch
    dev = cuda_dev
    main = torch.cuda.current_stream()
    dev.SetStream(main.cuda_stream)
    rng = np.random.default_rng(77)
    R, C = 257, 1000
    ta = torch.zeros(R, C, device="cuda", dtype=torch.float32)
    tb = torch.zeros(R, C, device="cuda", dtype=torch.float32)

    def w(t, dt):
        return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dt, owner=t)
    a, b = w(ta, dtypes.DN_F32), w(tb, dtypes.DN_F32)
    t1, t2, c = (Tensor.empty((R, C), dtypes.DN_F32, dev) for _ in range(3))
    ct = Tensor.empty((C, R), dtypes.DN_F32, dev)
    s1 = Tensor.empty((R,), dtypes.DN_F32, dev)
    mx = Tensor.empty((R,), dtypes.DN_F32, dev)
    am = Tensor.empty((R,), dtypes.DN_I64, dev)
    mx2 = Tensor.empty((R,), dtypes.DN_F32, dev)
    am2 = Tensor.empty((R,), dtypes.DN_I64, dev)
    _d = lambda t: t.Backend._d(t)

    def chain():
        t1.FillMultiply(a, b)
        t2.FillSin(a)
        c.FillAdd(t1, t2)
        s1.FillSumAxis(1, c)
        ct.FillAdd(a.T, b.T[:, 0:1].broadcastTo((C, R)))     # transposed + broadcast operands
        am._fill_axis("ArgMaxLastAxis", 1, c, True)
        mx.FillMaxAxis(1, c)
        dev.api.call("shard_minmax_arg_last_axis", None, 0, 1, _d(mx2), _d(am2), 0, _d(c))

    chain()                                   # every kernel loaded before the capture
    dev.Synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=side):
            dev.SetStream(side.cuda_stream)
            chain()
    finally:
        dev.SetStream(main.cuda_stream)
        dev.api.call("release_stream", side.cuda_stream)   # the library forgets the capture stream
    for rep in range(3):
        a_np = rng.uniform(-50, 50, size=(R, C)).astype(np.float32)
        b_np = rng.uniform(-50, 50, size=(R, C)).astype(np.float32)
        ta.copy_(torch.from_numpy(a_np))
        tb.copy_(torch.from_numpy(b_np))
        g.replay()
        torch.cuda.synchronize()
        ha, hb = HostTensor.ofNumpy(a_np), HostTensor.ofNumpy(b_np)
        hc = ha * hb + ha.sin()
        np.testing.assert_allclose(c.toNumpy(), hc.toNumpy(), rtol=1e-5, atol=1e-5)
        got_c = c.toNumpy()                  # the reductions are checked on the device's own c (exact)
        hc2 = HostTensor.ofNumpy(got_c)
        np.testing.assert_allclose(s1.toNumpy(), hc2.sumAxis(1).toNumpy(), rtol=1e-3)
        np.testing.assert_array_equal(am.toNumpy(), hc2.argMaxAxis(1).toNumpy())
        np.testing.assert_array_equal(mx.toNumpy(), hc2.maxAxis(1).toNumpy())
        np.testing.assert_array_equal(am2.toNumpy(), am.toNumpy())
        np.testing.assert_array_equal(mx2.toNumpy(), mx.toNumpy())
        want_ct = (ha.T + hb.T[:, 0:1].broadcastTo((C, R))).toNumpy()
        np.testing.assert_array_equal(ct.toNumpy(), want_ct)
