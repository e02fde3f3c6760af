"""The storage / transfer half of the boundary on a real device: native transfer of arbitrary views
(dn_transfer_h2d / dn_transfer_d2h; reference: CudaBackend.fs:206-270, round-trip tests CudaTests.fs:29-51),
registered host memory (CudaRegMem.fs:114-146), finalizer-safe release (dn_free_deferred; the reference frees from
the .NET finalizer thread, CudaBackend.fs:73-74) and the per-device function attributes a multi-device process needs."""
import ctypes as C
import threading

import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes, native
from deepnet_b200 import layout as TL
from deepnet_b200.backend import TensorStagingDevice, TensorStagingStorage
from helpers import ALL_DTYPES, rand_array

pytestmark = pytest.mark.gpu


def staging_view(arr: np.ndarray, view) -> Tensor:
    """A host staging tensor over `arr` (C-contiguous) with `view` applied: a strided HOST view."""
    st = TensorStagingStorage(arr.reshape(-1), TensorStagingDevice.Instance())
    return view(Tensor(TL.newC(arr.shape), st))


VIEWS = {
    "contiguous": lambda t: t,
    "transposed": lambda t: t.permuteAxes([2, 0, 1]),
    "sliced": lambda t: t[1:, 2:9, 1:4],
    "reversed": lambda t: t.reverseAxis(1),
    "inner strided": lambda t: t[:, :, 2],
    "broadcast source": lambda t: t[0:1].broadcastTo(t.Shape),
}


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I8, dtypes.DN_I16, dtypes.DN_I64, dtypes.DN_BOOL])
def test_transfer_of_strided_views_both_ways(cuda_dev, dtype):
    rng = np.random.default_rng(5)
    arr = rand_array(rng, (11, 13, 5), dtype)
    for hname, hview in VIEWS.items():
        for dname, dview in VIEWS.items():
            if dname == "broadcast source":
                continue  # a broadcast view cannot be a transfer TARGET
            src = staging_view(arr.copy(), hview)
            want = src.toNumpy()
            # host view -> device view of the same shape
            full = Tensor.zeros(arr.shape, dtype, cuda_dev)
            dv = dview(full)
            if dv.Shape != src.Shape:
                continue
            dv.TransferFrom(src)
            np.testing.assert_array_equal(dv.toNumpy(), want, err_msg=f"h2d {hname} -> {dname}")
            # device view -> strided host view (not a broadcast one)
            if hname == "broadcast source":
                continue
            back = np.zeros_like(arr)
            hv = staging_view(back, hview)
            hv.TransferFrom(dv)
            np.testing.assert_array_equal(hv.toNumpy(), want, err_msg=f"d2h {dname} -> {hname}")
            # elements of the host array outside the view stay untouched
            covered = staging_view(np.arange(arr.size, dtype=np.int64).reshape(arr.shape), hview).toNumpy().ravel()
            untouched = np.ones(arr.size, dtype=bool)
            untouched[covered] = False
            assert not back.ravel()[untouched].any(), f"d2h {dname} -> {hname} wrote outside the view"


def test_transfer_rank0_empty_and_large(cuda_dev):
    s = Tensor.ofNumpy(np.array(3.5, dtype=np.float64))
    assert CudaTensor.transfer(s).toNumpy() == 3.5
    e = Tensor.ofNumpy(np.zeros((0, 7), dtype=np.int32))
    assert CudaTensor.transfer(e).toNumpy().shape == (0, 7)
    rng = np.random.default_rng(6)
    big = rng.integers(-9, 9, size=(2048, 1031), dtype=np.int64)
    c = CudaTensor.ofNumpy(big)
    np.testing.assert_array_equal(c.T.toNumpy(), big.T)                      # strided device source
    hv = staging_view(np.zeros((1031, 2048), dtype=np.int64), lambda t: t.T)  # strided host target
    hv.TransferFrom(c)
    np.testing.assert_array_equal(hv.toNumpy(), big)


def test_registered_host_memory(cuda_dev):
    """CudaRegMem.register: an existing (page-aligned) host array is page-locked, transfers from it are DMA."""
    api = cuda_dev.api
    raw = np.zeros(1 << 22, dtype=np.uint8)
    off = (-raw.ctypes.data) % 4096
    buf = raw[off:off + (1 << 21)]
    api.call("host_register", buf.ctypes.data, buf.nbytes)
    api.call("host_register", buf.ctypes.data, buf.nbytes)       # registering twice is fine (CudaRegMem.fs:123-127)
    try:
        buf[:] = np.arange(buf.size, dtype=np.uint8)
        d = Tensor.empty((buf.size,), dtypes.DN_U8, cuda_dev)
        api.call("memcpy_h2d", d.Storage.BasePtr(), buf.ctypes.data, buf.nbytes)
        np.testing.assert_array_equal(d.toNumpy(), buf)
    finally:
        api.call("host_unregister", buf.ctypes.data)
    api.call("host_unregister", buf.ctypes.data)                 # and so is unregistering twice
    with pytest.raises(ValueError):                              # CannotCudaRegisterMemoryException analogue
        api.call("host_register", 0, 4096)


def test_deferred_free_from_a_foreign_thread(cuda_dev):
    """A 'finalizer' thread that never touched the device releases storages whose last consumer is still queued:
    dn_free_deferred only queues; the owner thread's next dn_alloc / dn_sync releases behind a stream fence. The
    results of the kernels that were in flight must be intact."""
    api = cuda_dev.api
    n = 1 << 24
    rng = np.random.default_rng(9)
    a_np = rng.integers(-100, 100, size=n, dtype=np.int64)
    a = CudaTensor.ofNumpy(a_np)
    outs = []
    for rep in range(8):
        p = C.c_void_p()
        api.call("alloc", n * 8, C.byref(p))
        tmp = CudaTensor.usingPtr(p.value, (n,), dtypes.DN_I64)
        tmp.FillAdd(a, a)                       # producer of tmp, queued
        out = tmp + a                           # consumer of tmp, queued
        th = threading.Thread(target=lambda q=p.value: api.call("free_deferred", q))
        th.start()
        th.join()
        outs.append(out)
        # allocations right after the deferred free must not receive tmp's block while `out` is being computed
        junk = [Tensor.filled((n,), rep, dtypes.DN_I64, cuda_dev) for _ in range(2)]
        del junk
    cuda_dev.Synchronize()
    for out in outs:
        np.testing.assert_array_equal(out.toNumpy(), 3 * a_np)
    with pytest.raises(ValueError):
        api.call("free_deferred", a_np.ctypes.data)   # not a device allocation


def test_free_is_ordered_after_other_streams(cuda_dev):
    """a storage freed on the thread's stream while its consumer runs on ANOTHER stream the library
    knows must not be recycled under that consumer."""
    import torch
    api = cuda_dev.api
    s2 = torch.cuda.Stream()
    n = 1 << 24
    a_np = np.arange(n, dtype=np.int64)
    a = CudaTensor.ofNumpy(a_np)
    cuda_dev.Synchronize()
    try:
        for rep in range(4):
            cuda_dev.SetStream(s2.cuda_stream)
            tmp = a + a                          # on s2
            out = tmp + a                        # on s2, reads tmp
            cuda_dev.SetStream(0)
            del tmp                              # freed on the default stream while s2 may still be reading it
            junk = Tensor.filled((n,), -1, dtypes.DN_I64, cuda_dev)   # default stream: would overwrite a recycled block
            cuda_dev.SetStream(s2.cuda_stream)
            cuda_dev.Synchronize()
            cuda_dev.SetStream(0)
            cuda_dev.Synchronize()
            np.testing.assert_array_equal(out.toNumpy(), 3 * a_np)
            del junk, out
    finally:
        cuda_dev.SetStream(0)
        api.call("release_stream", s2.cuda_stream)


def test_gemm_on_every_visible_device(cuda_dev):
    """VERDICT r01 / ADVICE: the > 48 KB shared-memory opt-in of the tcgen05 kernel is per device; a process that
    drives several devices (dn_set_device) must be able to launch it on each."""
    api = cuda_dev.api
    n = C.c_int32()
    api.call("device_count", C.byref(n))
    rng = np.random.default_rng(3)
    a_np = rng.uniform(-1, 1, size=(256, 512)).astype(np.float32)
    b_np = rng.uniform(-1, 1, size=(512, 384)).astype(np.float32)
    want = a_np.astype(np.float64) @ b_np.astype(np.float64)
    try:
        for d in range(n.value):
            api.call("init", d)
            got = (CudaTensor.ofNumpy(a_np) @ CudaTensor.ofNumpy(b_np)).toNumpy()
            assert np.linalg.norm(got - want) <= 1e-2 * np.linalg.norm(want), f"device {d}"
    finally:
        api.call("set_device", 0)
