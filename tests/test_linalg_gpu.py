"""BatchedInvert parity (TensorBackend.fs:142): the CUDA Gauss-Jordan kernel against the HostTensor oracle (LU with
partial pivoting = the host's LAPACK getrf + getri), plus the reference's own invert tests restated
(Tensor.Test/BaseTests.fs:162-211: diagonal, random 4x4, batch [2,4,3,3], singular matrix must fail).
Tolerance: rel 1e-4 (f32) / 1e-11 (f64) of the largest entry of the inverse, on well-conditioned matrices."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, SingularMatrixException, Tensor, dtypes
from helpers import pair

pytestmark = pytest.mark.gpu


def well_conditioned(rng, shape, npdt):
    n = shape[-1]
    m = rng.uniform(-1, 1, size=shape)
    return (m + np.eye(n) * (0.5 * n + 1)).astype(npdt)


def close(got, want, npdt, what):
    tol = (1e-4 if npdt == np.float32 else 1e-11) * max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    assert got.shape == want.shape and err <= tol, f"{what}: max err {err} > {tol}"


@pytest.mark.parametrize("npdt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 16, 33, 64, 100, 160, 230])
def test_invert_matches_host(cuda_dev, npdt, n):
    """n <= 158 (f64) / 224 (f32) runs in shared memory, larger matrices in global memory."""
    rng = np.random.default_rng(100 + n)
    m = well_conditioned(rng, (3, n, n), npdt)
    h, c = pair(m)
    hi, ci = Tensor.invert(h), Tensor.invert(c)
    close(ci.toNumpy(), hi.toNumpy(), npdt, f"invert n={n}")
    close(ci.toNumpy(), np.linalg.inv(m.astype(np.float64)), npdt, f"invert vs numpy n={n}")


@pytest.mark.parametrize("npdt", [np.float32, np.float64])
def test_invert_pivoting_views_in_place(cuda_dev, npdt):
    rng = np.random.default_rng(7)
    # zero diagonal: every step needs a row exchange
    m = rng.uniform(-1, 1, size=(5, 12, 12)).astype(npdt)
    for i in range(12):
        m[:, i, i] = 0
    h, c = pair(m)
    close(Tensor.invert(c).toNumpy(), Tensor.invert(h).toNumpy(), npdt, "zero diagonal")
    # transposed source view, sliced (strided) target
    m2 = well_conditioned(rng, (2, 4, 9, 9), npdt)
    h2, c2 = pair(m2)
    close(Tensor.invert(c2.swapDim(2, 3)).toNumpy(), Tensor.invert(h2.swapDim(2, 3)).toNumpy(), npdt, "transposed source")
    ht, ct = pair(np.zeros((2, 4, 12, 11), dtype=npdt))
    ht[:, :, 1:10, 2:11].FillInvert(h2)
    ct[:, :, 1:10, 2:11].FillInvert(c2)
    close(ct.toNumpy(), ht.toNumpy(), npdt, "strided target")
    # in place (target == source) and a broadcast source (one matrix inverted into every batch slot)
    c2.FillInvert(c2)
    close(c2.toNumpy(), Tensor.invert(h2).toNumpy(), npdt, "in place")
    one = well_conditioned(rng, (6, 6), npdt)
    h1, c1 = pair(one)
    hb, cb = pair(np.zeros((3, 6, 6), dtype=npdt))
    hb.FillInvert(h1)
    cb.FillInvert(c1)
    close(cb.toNumpy(), hb.toNumpy(), npdt, "broadcast source")
    # large strided matrices take the global-memory path through a dense scratch copy
    big = well_conditioned(rng, (300, 300), npdt)
    hB, cB = pair(big)
    close(Tensor.invert(cB.T).toNumpy(), Tensor.invert(hB.T).toNumpy(), npdt, "large transposed")


def test_reference_invert_tests(cuda_dev):
    """Tensor.Test/BaseTests.fs:162-211 on the CUDA device."""
    dm = CudaTensor.ofNumpy(np.diag([1.0, 2.0, 3.0]))
    np.testing.assert_allclose(Tensor.invert(Tensor.invert(dm)).toNumpy(), dm.toNumpy(), rtol=1e-12, atol=1e-12)
    rng = np.random.default_rng(123)
    for shape in [(4, 4), (2, 4, 3, 3)]:
        m = CudaTensor.ofNumpy(rng.uniform(-1, 1, size=shape))
        np.testing.assert_allclose(Tensor.invert(Tensor.invert(m)).toNumpy(), m.toNumpy(), rtol=1e-8, atol=1e-8)
    singular = CudaTensor.ofNumpy(np.array([[1.0, 0.0, 0.0], [1.0, 2.0, 0.0], [1.0, 0.0, 0.0]]))
    with pytest.raises(SingularMatrixException):
        Tensor.invert(singular)
    # one singular matrix inside a batch fails the whole call, like LAPACK info > 0 in the host loop
    batch = np.stack([np.eye(3), np.array([[1.0, 0, 0], [1, 2, 0], [1, 0, 0]]), np.eye(3)])
    with pytest.raises(SingularMatrixException):
        Tensor.invert(CudaTensor.ofNumpy(batch))
    with pytest.raises(Exception):
        Tensor.invert(CudaTensor.ofNumpy(np.arange(6, dtype=np.int32).reshape(1, 2, 3)))


def test_invert_many_small_matrices(cuda_dev):
    rng = np.random.default_rng(9)
    m = well_conditioned(rng, (4096, 8, 8), np.float32)
    got = Tensor.invert(CudaTensor.ofNumpy(m)).toNumpy()
    close(got, np.linalg.inv(m.astype(np.float64)), np.float32, "4096 x 8x8")
