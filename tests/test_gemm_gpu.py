"""MatMatDot / BatchedMatMatDot / MatVecDot / VecVecDot parity against the fp64-accumulated oracle
(host: MKL sgemm/dgemm, HostBackend.fs:463-546; reference tests: CudaTests.fs:54-75 `h .* i` GPU == host,
BaseTests.fs:85-96 batched == per-sample loop). float32 runs on tcgen05 with TF32 inputs: tolerance rel 1e-2
(north_star), measured norm-wise per row so that cancelling dot products are judged on the scale of their terms."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes
from helpers import pair, rand_array

pytestmark = pytest.mark.gpu


def check_mm(hc: Tensor, cc: Tensor, a: np.ndarray, b: np.ndarray, dtype: int, what: str):
    h, c = hc.toNumpy().astype(np.float64), cc.toNumpy().astype(np.float64)
    assert h.shape == c.shape, what
    # scale of the terms of each dot product: |a| @ |b|
    scale = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    tol = (2e-3 if dtype == dtypes.DN_F32 else 1e-13) * scale + 1e-30
    err = np.abs(h - c)
    assert (err <= tol).all(), f"{what}: max err/scale {np.max(err / (scale + 1e-30)):.3e}"
    if dtype == dtypes.DN_F32 and h.size:
        rel = np.linalg.norm(h - c) / (np.linalg.norm(h) + 1e-30)
        assert rel <= 1e-2, f"{what}: norm-wise rel err {rel:.3e} > 1e-2"


SHAPES = [(128, 256, 32), (128, 256, 64), (256, 512, 128), (5, 3, 3), (1, 1, 1), (130, 70, 45), (257, 300, 1000),
          (1000, 10, 4096), (10, 4096, 1000), (512, 784, 64), (300, 33, 7), (64, 2048, 2048)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_mat_mat_dot_row_major(cuda_dev, M, N, K, dtype):
    rng = np.random.default_rng(41)
    a, b = rand_array(rng, (M, K), dtype, -1, 1), rand_array(rng, (K, N), dtype, -1, 1)
    (ha, ca), (hb, cb) = pair(a), pair(b)
    check_mm(ha @ hb, ca @ cb, a, b, dtype, f"{M}x{K} . {K}x{N}")


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_mat_mat_dot_layouts(cuda_dev, dtype):
    """Every operand layout the MLP of config 5 produces: X·W.T, dY·W, dY.T·X, plus sliced / offset views and a
    non-row-major target."""
    rng = np.random.default_rng(42)
    M, N, K = 192, 160, 136
    a, bt = rand_array(rng, (M, K), dtype, -1, 1), rand_array(rng, (N, K), dtype, -1, 1)
    (ha, ca), (hbt, cbt) = pair(a), pair(bt)
    check_mm(ha @ hbt.T, ca @ cbt.T, a, bt.T, dtype, "A . B^T (both K-major)")
    at = rand_array(rng, (K, M), dtype, -1, 1)
    b = rand_array(rng, (K, N), dtype, -1, 1)
    (hat, cat), (hb, cb) = pair(at), pair(b)
    check_mm(hat.T @ hb, cat.T @ cb, at.T, b, dtype, "A^T . B (both MN-major)")
    check_mm(ha @ hb, ca @ cb, a, b, dtype, "A . B")
    check_mm(hat.T @ hbt.T, cat.T @ cbt.T, at.T, bt.T, dtype, "A^T . B^T")
    # sliced views with odd offsets and pitches
    check_mm(ha[3:131, 5:70] @ hb[5:70, 1:98], ca[3:131, 5:70] @ cb[5:70, 1:98], a[3:131, 5:70], b[5:70, 1:98], dtype,
             "sliced")
    # column-major target
    ht = Tensor.empty((M, N), dtype, ha.Dev, order="F")
    ct = Tensor.empty((M, N), dtype, ca.Dev, order="F")
    ht.FillDot(ha, hb)
    ct.FillDot(ca, cb)
    check_mm(ht, ct, a, b, dtype, "column-major target")


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_batched_and_broadcast(cuda_dev, dtype):
    """BaseTests.fs:85-114: batched matmul equals the per-sample loop; batch dims broadcast."""
    rng = np.random.default_rng(43)
    a, b = rand_array(rng, (3, 4, 70, 50), dtype, -1, 1), rand_array(rng, (3, 4, 50, 90), dtype, -1, 1)
    (ha, ca), (hb, cb) = pair(a), pair(b)
    hc, cc = ha @ hb, ca @ cb
    assert cc.Shape == (3, 4, 70, 90)
    for i in range(3):
        for j in range(4):
            check_mm(hc[i, j], cc[i, j], a[i, j], b[i, j], dtype, f"batch {i},{j}")
            check_mm(hc[i, j], ca[i, j] @ cb[i, j], a[i, j], b[i, j], dtype, f"batch {i},{j} vs 2-D call")
    b1 = rand_array(rng, (1, 1, 50, 90), dtype, -1, 1)
    hb1, cb1 = pair(b1)
    hc, cc = ha @ hb1, ca @ cb1
    check_mm(hc[2, 3], cc[2, 3], a[2, 3], b1[0, 0], dtype, "broadcast batch")


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_batched_many_small_and_large(cuda_dev, dtype):
    """Batches of small matrices run as ONE launch of the SIMT kernel (exact fp32 / fp64 accumulation): compared
    with numpy's batched matmul at fp32/fp64 rounding, incl. strided views, a broadcast batch operand and more
    than 65535 matrices; batches of large f32 matrices keep the per-element tcgen05 path."""
    rng = np.random.default_rng(46)
    npdt = dtypes.to_numpy(dtype)
    rtol = 2e-5 if dtype == dtypes.DN_F32 else 1e-12

    def check(ca, cb, a, b, what):
        want = np.matmul(a.astype(np.float64), b.astype(np.float64))
        scale = np.matmul(np.abs(a).astype(np.float64), np.abs(b).astype(np.float64)) + 1e-30
        got = (ca @ cb).toNumpy().astype(np.float64)
        assert got.shape == want.shape and (np.abs(got - want) <= rtol * scale).all(), what

    a, b = rand_array(rng, (1000, 8, 5), dtype, -1, 1), rand_array(rng, (1000, 5, 9), dtype, -1, 1)
    check(CudaTensor.ofNumpy(a), CudaTensor.ofNumpy(b), a, b, "1000 x [8,5].[5,9]")
    a, b = rand_array(rng, (7, 3, 65, 70), dtype, -1, 1), rand_array(rng, (1, 3, 70, 130), dtype, -1, 1)
    ca, cb = CudaTensor.ofNumpy(a), CudaTensor.ofNumpy(b)
    check(ca, cb.broadcastTo((7, 3, 70, 130)), a, np.broadcast_to(b, (7, 3, 70, 130)), "broadcast batch dim")
    # transposed matrices inside the batch (K-major vs MN-major operands) and a sliced batch
    at = rand_array(rng, (5, 40, 33), dtype, -1, 1)
    bt = rand_array(rng, (5, 21, 40), dtype, -1, 1)
    check(CudaTensor.ofNumpy(at).swapDim(1, 2)[1:4], CudaTensor.ofNumpy(bt).swapDim(1, 2)[1:4],
          np.swapaxes(at, 1, 2)[1:4], np.swapaxes(bt, 1, 2)[1:4], "transposed views")
    a, b = rand_array(rng, (70000, 2, 3), dtype, -1, 1), rand_array(rng, (70000, 3, 2), dtype, -1, 1)
    check(CudaTensor.ofNumpy(a), CudaTensor.ofNumpy(b), a, b, "70000 matrices (two launches)")
    if dtype == dtypes.DN_F32:  # large matrices per batch element: tcgen05 per element, tf32 tolerance
        a, b = rand_array(rng, (2, 600, 520), dtype, -1, 1), rand_array(rng, (2, 520, 560), dtype, -1, 1)
        (ha, ca), (hb, cb) = pair(a), pair(b)
        hc, cc = ha @ hb, ca @ cb
        for i in range(2):
            check_mm(hc[i], cc[i], a[i], b[i], dtype, f"large batch element {i}")


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_vec_dots(cuda_dev, dtype):
    rng = np.random.default_rng(44)
    a, x, y = rand_array(rng, (300, 1000), dtype, -1, 1), rand_array(rng, (1000,), dtype, -1, 1), \
        rand_array(rng, (1000,), dtype, -1, 1)
    (ha, ca), (hx, cx), (hy, cy) = pair(a), pair(x), pair(y)
    tol = 1e-4 if dtype == dtypes.DN_F32 else 1e-12
    s_h, s_c = float((hx @ hy).Value), float((cx @ cy).Value)
    assert abs(s_h - s_c) <= tol * float(np.abs(x.astype(np.float64)) @ np.abs(y.astype(np.float64)))
    for (hm, cm, an) in [(ha, ca, a), (ha[10:200, :], ca[10:200, :], a[10:200, :])]:
        hv, cv = (hm @ hx).toNumpy().astype(np.float64), (cm @ cx).toNumpy().astype(np.float64)
        scale = np.abs(an.astype(np.float64)) @ np.abs(x.astype(np.float64))
        assert (np.abs(hv - cv) <= tol * scale).all()
    at = rand_array(rng, (1000, 300), dtype, -1, 1)
    hat, cat = pair(at)
    hv, cv = (hat.T @ hx).toNumpy().astype(np.float64), (cat.T @ cx).toNumpy().astype(np.float64)
    assert (np.abs(hv - cv) <= tol * (np.abs(at.T.astype(np.float64)) @ np.abs(x.astype(np.float64)))).all()


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_vec_dots_layouts(cuda_dev, dtype):
    """Every kernel variant of VecVecDot / MatVecDot: 128-bit path and its tail, unaligned bases, strided and
    reversed operands, transposed matrices with and without K-slices, strided targets."""
    rng = np.random.default_rng(45)
    tol = 1e-4 if dtype == dtypes.DN_F32 else 1e-12
    f64 = lambda v: v.astype(np.float64)
    for n in (1, 3, 4, 1027, 70001):
        x, y = rand_array(rng, (n + 3,), dtype, -1, 1), rand_array(rng, (n + 3,), dtype, -1, 1)
        (hx, cx), (hy, cy) = pair(x), pair(y)
        for sl_x, sl_y in [((slice(0, n),), (slice(0, n),)), ((slice(1, n + 1),), (slice(0, n),)),
                           ((slice(3, n + 3),), (slice(2, n + 2),))]:
            s_h, s_c = float((hx[sl_x] @ hy[sl_y]).Value), float((cx[sl_x] @ cy[sl_y]).Value)
            assert abs(s_h - s_c) <= tol * float(np.abs(f64(x[sl_x])) @ np.abs(f64(y[sl_y]))) + 1e-30
    x, y = rand_array(rng, (2000, 2), dtype, -1, 1), rand_array(rng, (2000,), dtype, -1, 1)
    (hx, cx), (hy, cy) = pair(x), pair(y)
    s_h, s_c = float((hx[:, 0] @ hy.reverseAxis(0)).Value), float((cx[:, 0] @ cy.reverseAxis(0)).Value)
    assert abs(s_h - s_c) <= tol * float(np.abs(f64(x[:, 0])) @ np.abs(f64(y[::-1])))
    for (m, k) in [(5, 7), (64, 4096), (33, 1026), (2048, 300), (1024, 5000)]:
        a, x2 = rand_array(rng, (m, k), dtype, -1, 1), rand_array(rng, (k, 2), dtype, -1, 1)
        x = np.ascontiguousarray(x2.reshape(-1))
        at = np.ascontiguousarray(a.T)
        (ha, ca), (hat, cat), (hx, cx), (hx2, cx2) = pair(a), pair(at), pair(x), pair(x2)
        scale = np.abs(f64(a)) @ np.abs(f64(x[:k])) + 1e-30
        for hm, cm in [(ha, ca), (hat.T, cat.T)]:
            for hv_, cv_ in [(hx[0:k], cx[0:k])]:
                hv, cv = f64((hm @ hv_).toNumpy()), f64((cm @ cv_).toNumpy())
                assert (np.abs(hv - cv) <= tol * scale).all(), (m, k)
            hv, cv = f64((hm @ hx2[:, 0]).toNumpy()), f64((cm @ cx2[:, 0]).toNumpy())
            assert (np.abs(hv - cv) <= tol * (np.abs(f64(a)) @ np.abs(f64(x2[:, 0])) + 1e-30)).all(), (m, k, "strided x")
        if m > 8:  # row / column sub-views: unaligned rows, unaligned leading column
            for hm, cm, an in [(ha[1:m - 2, 1:k], ca[1:m - 2, 1:k], a[1:m - 2, 1:k]),
                               (hat.T[1:m - 2, 1:k], cat.T[1:m - 2, 1:k], a[1:m - 2, 1:k]),
                               (ha.reverseAxis(0), ca.reverseAxis(0), a[::-1, :])]:
                kk = an.shape[1]
                hv, cv = f64((hm @ hx[0:kk]).toNumpy()), f64((cm @ cx[0:kk]).toNumpy())
                assert (np.abs(hv - cv) <= tol * (np.abs(f64(an)) @ np.abs(f64(x[:kk])) + 1e-30)).all(), (m, k, "sub-view")
        # strided target
        for cm in (ca, cat.T):
            tgt = CudaTensor.zeros((m, 2), dtype)
            tgt[:, 0].FillDot(cm, cx[0:k])
            got = f64(tgt.toNumpy())
            want = f64((ha @ hx[0:k]).toNumpy())
            assert (np.abs(got[:, 0] - want) <= tol * scale).all() and (got[:, 1] == 0).all()


def test_unsupported_types(cuda_dev):
    from deepnet_b200.native import NotSupportedException
    a = CudaTensor.zeros((4, 4), dtypes.DN_I32)
    with pytest.raises(NotSupportedException):
        a @ a


# ---------------------------------------------------------------------------------------------------------------
# precision modes (dn_set_math_mode)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture
def tf32_mode(cuda_dev):
    cuda_dev.SetMathMode("tf32")
    yield
    cuda_dev.SetMathMode("fp32")


def test_reference_single_matrix_dot(cuda_dev):
    """Tensor.Test/CudaTests.fs:52-62 `Single matrix dot`: h = init [5;3] (3i + j), i = 0.1 + identity 3,
    GPU result almostEqual the host's (absTol 1e-5, relTol 1e-5 — Tensor.almostEqual's defaults)."""
    h = np.array([[3.0 * i + j for j in range(3)] for i in range(5)], dtype=np.float32)
    m = (0.1 + np.eye(3)).astype(np.float32)
    (hh, ch), (hm, cm) = pair(h), pair(m)
    want, got = (hh @ hm).toNumpy(), (ch @ cm).toNumpy()
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)
    # the same product through the batched entry point gives the same bits (one precision inside the library)
    got_b = (ch.reshape((1, 5, 3)) @ cm.reshape((1, 3, 3))).toNumpy()[0]
    assert (got_b == got).all()


FP32_SHAPES = [(5, 3, 3), (130, 70, 45), (512, 512, 512), (1024, 1024, 1024), (777, 333, 1111), (2048, 512, 4096),
               (8192, 10, 4096), (300, 4096, 784)]


def split_tol(K: int) -> float:
    """Norm-wise relative error budget of the 3xTF32 path: the tensor core truncates its fp32 accumulator after every
    MMA (3 * K/8 of them per output), a bias that grows linearly with K — measured 7e-6 at K = 1024."""
    return 2e-6 + 1.2e-8 * K


@pytest.mark.parametrize("M,N,K", FP32_SHAPES)
def test_default_math_mode(cuda_dev, M, N, K):
    """Default precision: products below 2^27 multiply-adds are exact fp32 (SIMT kernel, norm-wise 1e-6), larger ones
    3xTF32 on the tensor cores (50-100x closer to the fp64 result than a single tf32 pass)."""
    rng = np.random.default_rng(47)
    a, b = rand_array(rng, (M, K), dtypes.DN_F32, -1, 1), rand_array(rng, (K, N), dtypes.DN_F32, -1, 1)
    got = (CudaTensor.ofNumpy(a) @ CudaTensor.ofNumpy(b)).toNumpy().astype(np.float64)
    want = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    err = np.abs(got - want)
    exact_path = M * N * K < (1 << 27)
    assert (err <= (1e-6 if exact_path else 1e-4) * scale + 1e-30).all(), f"max err/scale {np.max(err / (scale + 1e-30)):.3e}"
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel <= (1e-6 if exact_path else split_tol(K)), rel


@pytest.mark.parametrize("M,N,K", [(1024, 1024, 1024), (2048, 512, 4096)])
def test_strict_math_mode_is_exact_fp32(cuda_dev, M, N, K):
    rng = np.random.default_rng(50)
    a, b = rand_array(rng, (M, K), dtypes.DN_F32, -1, 1), rand_array(rng, (K, N), dtypes.DN_F32, -1, 1)
    want = a.astype(np.float64) @ b.astype(np.float64)
    cuda_dev.SetMathMode("strict")
    try:
        got = (CudaTensor.ofNumpy(a) @ CudaTensor.ofNumpy(b)).toNumpy().astype(np.float64)
    finally:
        cuda_dev.SetMathMode("fp32")
    assert np.linalg.norm(got - want) <= 3e-6 * np.linalg.norm(want)   # sqrt(K) * 2^-24 accumulation, K up to 4096


def test_fp32_mode_layouts_and_specials(cuda_dev):
    """3xTF32 path with transposed / sliced operands, a strided target, and non-finite inputs (the split must not
    turn an infinity into a NaN: lo = 0 for non-finite elements)."""
    rng = np.random.default_rng(48)
    M, N, K = 640, 384, 1024
    a, bt = rand_array(rng, (M, K), dtypes.DN_F32, -1, 1), rand_array(rng, (N, K), dtype=dtypes.DN_F32, lo=-1, hi=1)
    at = np.ascontiguousarray(a.T)
    ca, cbt, cat = CudaTensor.ofNumpy(a), CudaTensor.ofNumpy(bt), CudaTensor.ofNumpy(at)
    want = a.astype(np.float64) @ bt.T.astype(np.float64)
    for what, got in (("A . B^T", ca @ cbt.T), ("A^T^T . B^T", cat.T @ cbt.T)):
        g = got.toNumpy().astype(np.float64)
        assert np.linalg.norm(g - want) <= split_tol(K) * np.linalg.norm(want), what
    sa, sb = ca[3:, 5:901], cbt[1:, 5:901]
    g = (sa @ sb.T).toNumpy().astype(np.float64)
    w = a[3:, 5:901].astype(np.float64) @ bt[1:, 5:901].T.astype(np.float64)
    assert np.linalg.norm(g - w) <= split_tol(K) * np.linalg.norm(w), "sliced"
    tgt = Tensor.empty((M, N), dtypes.DN_F32, cuda_dev, order="F")
    tgt.FillDot(ca, cbt.T)
    assert np.linalg.norm(tgt.toNumpy().astype(np.float64) - want) <= split_tol(K) * np.linalg.norm(want), "column-major target"
    a2 = a.copy()
    a2[7, 11] = np.inf
    a2[9, 13] = np.nan
    g = (CudaTensor.ofNumpy(a2) @ cbt.T).toNumpy()
    w = a2.astype(np.float64) @ bt.T.astype(np.float64)
    assert (np.isnan(g) == np.isnan(w)).all() and (np.isinf(g) == np.isinf(w)).all()
    assert (np.sign(g[np.isinf(g)]) == np.sign(w[np.isinf(w)])).all()


def test_tf32_mode_is_opt_in(cuda_dev, tf32_mode):
    """DN_MATH_TF32: one tcgen05 pass with tf32 inputs — inside rel 1e-2 (north_star) and measurably coarser than
    the default mode on the same data (which is how the test knows the switch did something)."""
    rng = np.random.default_rng(49)
    a, b = rand_array(rng, (1024, 1024), dtypes.DN_F32, -1, 1), rand_array(rng, (1024, 1024), dtypes.DN_F32, -1, 1)
    want = a.astype(np.float64) @ b.astype(np.float64)
    got = (CudaTensor.ofNumpy(a) @ CudaTensor.ofNumpy(b)).toNumpy().astype(np.float64)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert 1e-4 < rel <= 1e-2, rel


TF32_SHAPES = [(512, 256, 64), (512, 512, 32), (1000, 300, 777), (1024, 1024, 1024), (777, 1111, 333), (8192, 4096, 784),
               (2048, 4096, 4096), (513, 257, 100)]


@pytest.mark.parametrize("M,N,K", TF32_SHAPES)
def test_tf32_mode_shapes(cuda_dev, tf32_mode, M, N, K):
    """DN_MATH_TF32 on shapes that take the CTA-pair kernel (M >= 512, N >= 256: 256 x 256 tiles, cta_group::2) with
    full, partial and odd tiles; rel 1e-2 (north_star), per element on the scale of the dot product's terms."""
    rng = np.random.default_rng(53)
    a, b = rand_array(rng, (M, K), dtypes.DN_F32, -1, 1), rand_array(rng, (K, N), dtypes.DN_F32, -1, 1)
    (ha, ca), (hb, cb) = pair(a), pair(b)
    check_mm(ha @ hb, ca @ cb, a, b, dtypes.DN_F32, f"tf32 {M}x{K} . {K}x{N}")


def test_tf32_mode_layouts(cuda_dev, tf32_mode):
    """Every operand layout of the MLP step through the CTA-pair kernel: X.W^T (K-major both), dY.W (B N-major),
    dY^T.X (A M-major), both transposed, sliced views (repacked), a column-major target."""
    rng = np.random.default_rng(54)
    M, N, K = 768, 512, 264
    a, bt = rand_array(rng, (M, K), dtypes.DN_F32, -1, 1), rand_array(rng, (N, K), dtypes.DN_F32, -1, 1)
    (ha, ca), (hbt, cbt) = pair(a), pair(bt)
    check_mm(ha @ hbt.T, ca @ cbt.T, a, bt.T, dtypes.DN_F32, "A . B^T")
    at, b = rand_array(rng, (K, M), dtypes.DN_F32, -1, 1), rand_array(rng, (K, N), dtypes.DN_F32, -1, 1)
    (hat, cat), (hb, cb) = pair(at), pair(b)
    check_mm(hat.T @ hb, cat.T @ cb, at.T, b, dtypes.DN_F32, "A^T . B")
    check_mm(ha @ hb, ca @ cb, a, b, dtypes.DN_F32, "A . B")
    check_mm(hat.T @ hbt.T, cat.T @ cbt.T, at.T, bt.T, dtypes.DN_F32, "A^T . B^T")
    check_mm(ha[3:700, 5:200] @ hb[5:200, 1:400], ca[3:700, 5:200] @ cb[5:200, 1:400], a[3:700, 5:200], b[5:200, 1:400],
             dtypes.DN_F32, "sliced")
    ht = Tensor.empty((M, N), dtypes.DN_F32, ha.Dev, order="F")
    ct = Tensor.empty((M, N), dtypes.DN_F32, ca.Dev, order="F")
    ht.FillDot(ha, hb)
    ct.FillDot(ca, cb)
    check_mm(ht, ct, a, b, dtypes.DN_F32, "column-major target")


def test_tf32_batched_one_launch(cuda_dev, tf32_mode):
    """BatchedMatMatDot with LARGE batch elements in tf32 mode is one persistent launch of the CTA-pair kernel over
    3-D tensor maps (the reference: one cublasSgemmBatched call, CudaBackend.fs:426-449): plain batches, a shared
    (broadcast) right operand, two batch dims, a transposed operand — each against the per-element 2-D call."""
    rng = np.random.default_rng(55)
    a = rand_array(rng, (3, 2, 512, 520), dtypes.DN_F32, -1, 1)      # 2^27.02 multiply-adds per element
    b = rand_array(rng, (3, 2, 520, 512), dtypes.DN_F32, -1, 1)
    (ha, ca), (hb, cb) = pair(a), pair(b)
    n0 = cuda_dev.LaunchCount()
    cc = ca @ cb
    assert cuda_dev.LaunchCount() - n0 == 1, "one launch for the whole batch"
    hc = ha @ hb
    for i in range(3):
        for j in range(2):
            check_mm(hc[i, j], cc[i, j], a[i, j], b[i, j], dtypes.DN_F32, f"batch {i},{j}")
    b1 = rand_array(rng, (1, 1, 520, 512), dtypes.DN_F32, -1, 1)
    hb1, cb1 = pair(b1)
    n0 = cuda_dev.LaunchCount()
    cc = ca @ cb1.broadcastTo((3, 2, 520, 512))
    assert cuda_dev.LaunchCount() - n0 == 1
    hc = ha @ hb1.broadcastTo((3, 2, 520, 512))
    check_mm(hc[2, 1], cc[2, 1], a[2, 1], b1[0, 0], dtypes.DN_F32, "shared right operand")
    check_mm(hc[0, 0], cc[0, 0], a[0, 0], b1[0, 0], dtypes.DN_F32, "shared right operand [0,0]")
    bt = rand_array(rng, (3, 2, 512, 520), dtypes.DN_F32, -1, 1)
    hbt, cbt = pair(bt)
    cc = ca @ cbt.swapDim(2, 3)
    hc = ha @ hbt.swapDim(2, 3)
    check_mm(hc[1, 1], cc[1, 1], a[1, 1], bt[1, 1].T, dtypes.DN_F32, "transposed right operand")


def test_one_cta_kernel_stays_covered():
    """Large tf32 products now take the CTA-pair kernel; DN_GEMM_2CTA=0 (test hook, read once per process, hence the child
    process) sends the tf32 cases of this file through the one-CTA 128 x 256 kernel again."""
    import os
    import subprocess
    import sys
    if os.environ.get("DN_GEMM_2CTA") == "0":
        pytest.skip("already running under the hook")
    env = dict(os.environ, DN_GEMM_2CTA="0")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-k", "tf32 and not batched", "-q", "-m", "gpu", "-x"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
