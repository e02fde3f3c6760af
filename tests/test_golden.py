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
