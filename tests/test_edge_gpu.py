"""Edge cases of the boundary: maximum rank, zero-sized tensors, > 2^31-element tensors (the element-wise launcher
splits them into chunks; the host backend cannot even allocate them, HostBackend.fs:54-58, so these are checked
through size-independent properties), error conventions."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes
from deepnet_b200 import layout as TL
from deepnet_b200.native import NotSupportedException
from helpers import assert_same, pair, rand_array

pytestmark = pytest.mark.gpu


def test_rank_8_and_rank_9(cuda_dev):
    rng = np.random.default_rng(61)
    shape = (2, 3, 2, 2, 3, 2, 2, 3)
    ha, ca = pair(rand_array(rng, shape, dtypes.DN_I32))
    hb, cb = pair(rand_array(rng, shape, dtypes.DN_I32))
    perm = [7, 0, 3, 1, 6, 2, 5, 4]
    assert_same(ha.permuteAxes(perm) + hb.permuteAxes(perm), ca.permuteAxes(perm) + cb.permuteAxes(perm), dtypes.DN_I32,
                what="rank-8 permuted add")
    assert_same(ha.permuteAxes(perm).sumAxis(3), ca.permuteAxes(perm).sumAxis(3), dtypes.DN_I32, what="rank-8 sumAxis")
    assert_same(ha.gt(0).trueIdx(), ca.gt(0).trueIdx(), dtypes.DN_I64, what="rank-8 trueIdx")
    nine = Tensor(TL.newC((1,) * 9), ca.Storage)
    with pytest.raises(NotSupportedException):
        nine + nine


def test_zero_sized(cuda_dev, host_dev):
    for shape in [(0,), (5, 0), (0, 7), (3, 0, 4)]:
        h, c = Tensor.zeros(shape, dtypes.DN_F32, host_dev), Tensor.zeros(shape, dtypes.DN_F32, cuda_dev)
        assert_same(h + h, c + c, dtypes.DN_F32, what=f"add {shape}")
        assert_same(h.sin(), c.sin(), dtypes.DN_F32, what=f"sin {shape}")
        assert_same(h.gt(0.0).trueIdx(), c.gt(0.0).trueIdx(), dtypes.DN_I64, what=f"trueIdx {shape}")
        assert c.gt(0.0).countTrue() == 0
        for ax in range(len(shape)):
            assert_same(h.sumAxis(ax), c.sumAxis(ax), dtypes.DN_F32, what=f"sumAxis {ax} {shape}")
            assert_same(h.maxAxis(ax), c.maxAxis(ax), dtypes.DN_F32, what=f"maxAxis {ax} {shape}")
            assert_same(h.argMaxAxis(ax), c.argMaxAxis(ax), dtypes.DN_I64, what=f"argMaxAxis {ax} {shape}")
    # empty sum = 0, empty product = 1, empty max = MinValue (Tensor.fs:2282, 2326, 2416 remarks)
    e = Tensor.zeros((4, 0), dtypes.DN_F64, cuda_dev)
    assert e.sumAxis(1).toNumpy().tolist() == [0.0] * 4
    assert e.productAxis(1).toNumpy().tolist() == [1.0] * 4
    assert e.maxAxis(1).toNumpy().tolist() == [np.finfo(np.float64).min] * 4
    assert (e @ Tensor.zeros((0, 3), dtypes.DN_F64, cuda_dev)).toNumpy().tolist() == [[0.0] * 3] * 4


def test_more_than_2_pow_31_elements(cuda_dev):
    """uint8 tensor of 2^31 + 1000 elements: chunked launches must cover every element exactly once."""
    import torch
    n = (1 << 31) + 1000
    ta = torch.empty(n, dtype=torch.uint8, device="cuda")
    tb = torch.empty(n, dtype=torch.uint8, device="cuda")
    a = CudaTensor.usingPtr(ta.data_ptr(), (n,), dtypes.DN_U8, owner=ta)
    b = CudaTensor.usingPtr(tb.data_ptr(), (n,), dtypes.DN_U8, owner=tb)
    a.FillConst(3)
    assert int(ta.to(torch.int64).sum().item()) == 3 * n
    b.FillAdd(a, a)                      # contiguous vector path, chunked
    assert int(tb.min().item()) == 6 and int(tb.max().item()) == 6
    # a 2-D strided view spanning more than 2^31 elements: [2, 2^30 + 500] over the same memory, reversed rows
    a2 = a.reshape((2, n // 2)).reverseAxis(1)
    b2 = b.reshape((2, n // 2))
    ta[: n // 2] = 1
    ta[n // 2:] = 2
    b2.FillMultiply(a2, a2)              # scalar path with a negative stride, chunked along the outer dim
    assert int(tb[: n // 2].max().item()) == 1 and int(tb[n // 2:].min().item()) == 4
    del ta, tb


def test_error_conventions(cuda_dev):
    a = CudaTensor.zeros((4, 5), dtypes.DN_F32)
    b = CudaTensor.zeros((4, 6), dtypes.DN_F32)
    with pytest.raises(RuntimeError):           # InvalidOperationException: cannot broadcast
        a + b
    with pytest.raises((RuntimeError, ValueError)):
        a.Backend.Add(a, a, b)                  # the backend itself rejects mismatched shapes
    with pytest.raises(ValueError):
        a.Backend.Add(a, a, CudaTensor.zeros((4, 5), dtypes.DN_F64))   # operand types differ
    with pytest.raises(NotSupportedException):
        CudaTensor.zeros((3,), dtypes.DN_BOOL) + CudaTensor.zeros((3,), dtypes.DN_BOOL)
    with pytest.raises(RuntimeError):           # InvalidOperationException: [4, 5] is not a (batch of) square matrices
        a.Backend.BatchedInvert(a, a)
    with pytest.raises(NotSupportedException):  # unsupported on CUDA in the reference too (CudaBackend.fs:486-488)
        a.Backend.BatchedSVD(a, a, a)
    ii = CudaTensor.zeros((3, 3), dtypes.DN_I32)
    with pytest.raises(NotSupportedException):  # "only supported for floating point numbers" (HostBackend.fs:549-550)
        ii.Backend.BatchedInvert(ii, ii)
    bc = a[0:1, :].broadcastTo((4, 5))
    with pytest.raises(ValueError):             # a broadcast view cannot be a target
        bc.FillAdd(a, a)


def test_concurrent_threads_with_their_own_streams(cuda_dev):
    """SURVEY §8b "Threading": any thread may call, Cfg.Stream is thread-local (CudaCfg.fs:16,25-27). Four threads,
    each on its own stream, run element-wise, reduction, compaction, matmul and invert calls concurrently (ctypes
    releases the GIL during the native calls); every result must equal the single-threaded oracle's."""
    import threading
    import torch
    from deepnet_b200 import Tensor
    from oracle.host_tensor import HostTensor

    def work(seed):
        rng = np.random.default_rng(seed)
        f = rng.uniform(-50, 50, size=(300, 257)).astype(np.float32)
        i = rng.integers(-1000, 1000, size=(300, 257)).astype(np.int64)
        sq = (rng.uniform(-1, 1, size=(8, 12, 12)) + 8 * np.eye(12)).astype(np.float64)
        out = {}
        for name, mk in (("host", HostTensor.ofNumpy), ("cuda", CudaTensor.ofNumpy)):
            a, b, m = mk(f), mk(i), mk(sq)
            out[name] = ((a.T + a.T.reverseAxis(0)).toNumpy(), (b * 3 - b.T.T).sumAxis(0).toNumpy(),
                         b.M(b.gt(0)).toNumpy(), b.gt(500).trueIdx().toNumpy(), a.argMaxAxis(1).toNumpy(),
                         Tensor.invert(m).toNumpy(), (m @ m).toNumpy())
        return out

    results, errors = {}, []

    def runner(k):
        try:
            stream = torch.cuda.Stream()
            cuda_dev.SetStream(stream.cuda_stream)
            for rep in range(3):
                results[(k, rep)] = work(1000 + 10 * k + rep)
            cuda_dev.Synchronize()
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=runner, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 12
    for key, out in results.items():
        for h, c in zip(out["host"][:5], out["cuda"][:5]):
            assert h.shape == c.shape and (h == c).all(), key
        np.testing.assert_allclose(out["cuda"][5], out["host"][5], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["cuda"][6], out["host"][6], rtol=1e-10, atol=1e-10)


def test_reference_diag_and_trace_tests(cuda_dev):
    """Tensor.Test/BaseTests.fs:128-158 on the CUDA device: diagMat writes through a diagonal view (stride n+1),
    trace reduces over one; results equal the host's bit for bit (integers) / exactly (copied floats)."""
    from deepnet_b200 import Tensor
    from oracle.host_tensor import HostTensor
    rng = np.random.default_rng(77)
    v = rng.uniform(-5, 5, size=(4, 33)).astype(np.float32)
    i = rng.integers(-100, 100, size=(5, 7, 65)).astype(np.int64)
    for arr in (v, i):
        h, c = HostTensor.ofNumpy(arr), CudaTensor.ofNumpy(arr)
        hd, cd = Tensor.diagMat(h), Tensor.diagMat(c)
        assert cd.Shape == arr.shape + (arr.shape[-1],)
        assert np.array_equal(cd.toNumpy(), hd.toNumpy())
        assert np.array_equal(cd.diag().toNumpy(), arr)
        if arr.dtype == np.int64:
            assert np.array_equal(cd.trace().toNumpy(), c.sumAxis(arr.ndim - 1).toNumpy())
            assert np.array_equal(cd.trace().toNumpy(), hd.trace().toNumpy())
        else:
            np.testing.assert_allclose(cd.trace().toNumpy(), hd.trace().toNumpy(), rtol=1e-5, atol=1e-4)
