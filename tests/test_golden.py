"""Pins the CPU oracle — and the CUDA path — against the reference's own documented known answers: every
`<example>` with an expected value in Tensor/Tensor/Tensor.fs and the printed results in
Tensor.Docs/articles/Guide-*.md for the ops on the hot path. Transcribed into
tests/golden/doc_kats.json with file:line per case."""
import pytest

from golden_runner import check_case, load_cases, run_case

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_docs(case):
    from oracle.host_tensor import HostTensor
    check_case(case, run_case(case, HostTensor.ofNumpy))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES], ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_docs(cuda_dev, case):
    from deepnet_b200 import CudaTensor
    if case["op"] == "dot" and False:
        pytest.skip()
    check_case(case, run_case(case, CudaTensor.ofNumpy))


def test_oracle_invert_against_numpy():
    """The oracle's getrf/getri restatement (HostBackend.fs:548-577) against numpy's LAPACK on random batches, the
    reference's own invert tests (Tensor.Test/BaseTests.fs:162-211) and the singular-matrix error."""
    import numpy as np
    import pytest
    from deepnet_b200 import SingularMatrixException, Tensor
    from oracle.host_tensor import HostTensor
    rng = np.random.default_rng(123)
    for npdt, tol in ((np.float64, 1e-10), (np.float32, 2e-4)):
        for shape in [(4, 4), (2, 4, 3, 3), (3, 40, 40)]:
            m = (rng.uniform(-1, 1, size=shape) + np.eye(shape[-1]) * 2).astype(npdt)
            got = Tensor.invert(HostTensor.ofNumpy(m)).toNumpy()
            want = np.linalg.inv(m.astype(np.float64))
            assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max())
            back = Tensor.invert(Tensor.invert(HostTensor.ofNumpy(m))).toNumpy()
            assert np.abs(back - m).max() <= 10 * tol
    with pytest.raises(SingularMatrixException):
        Tensor.invert(HostTensor.ofNumpy(np.array([[1.0, 0.0, 0.0], [1.0, 2.0, 0.0], [1.0, 0.0, 0.0]])))


def test_reference_diag_and_trace_tests_on_oracle():
    """Tensor.Test/BaseTests.fs:128-158 (`Build and extract diagonal`, `Batched trace`) on the HostTensor oracle."""
    import numpy as np
    from deepnet_b200 import Tensor
    from oracle.host_tensor import HostTensor
    v = HostTensor.ofNumpy(np.array([1.0, 2.0, 3.0]))
    dm = Tensor.diagMat(v)
    assert np.array_equal(dm.toNumpy(), np.diag([1.0, 2.0, 3.0])) and np.array_equal(dm.diag().toNumpy(), v.toNumpy())
    vb = HostTensor.ofNumpy(np.array([[1, 2, 3], [4, 5, 6]], dtype=np.int32))
    dmb = Tensor.diagMat(vb)
    assert dmb.Shape == (2, 3, 3)
    assert np.array_equal(dmb.trace().toNumpy(), vb.sumAxis(1).toNumpy())
