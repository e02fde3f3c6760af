"""Config 5 parity: one LearnMnist-style MLP training step (tools/mlp_step.py) on CudaTensor vs the HostTensor
oracle at a CPU-sized configuration, tolerance rel 1e-2 (TF32 tensor-core GEMM, north_star), then the full-size
configuration (784-4096-4096-10, batch 8192) checked through size-independent properties: softmax rows sum to 1,
the loss decreases under SGD, and the loss equals the one recomputed from the returned predictions."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from mlp_step import init_params, synthetic_batch, train_step  # noqa: E402

from deepnet_b200 import CudaTensor  # noqa: E402
from oracle.host_tensor import HostTensor  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fp32", "tf32"])
def math_mode(request, cuda_dev):
    """Both precisions of float32 MatMatDot (dn_set_math_mode): the fp32-accurate default and the opt-in tf32 pass."""
    cuda_dev.SetMathMode(request.param)
    yield request.param
    cuda_dev.SetMathMode("fp32")


def test_mlp_step_matches_oracle(cuda_dev, math_mode):
    sizes, batch = (784, 192, 160, 10), 256
    rng = np.random.default_rng(51)
    p0 = init_params(rng, sizes)
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    out = {}
    for name, mk in (("host", HostTensor.ofNumpy), ("cuda", CudaTensor.ofNumpy)):
        params = [(mk(w.copy()), mk(b.copy())) for w, b in p0]
        x, t = mk(xn), mk(tn)
        losses = []
        for _ in range(2):
            loss, pred = train_step(x, t, params, 0.05)
            losses.append(float(loss.Value))
        out[name] = (losses, pred.toNumpy(), [(w.toNumpy(), b.toNumpy()) for w, b in params])
    (hl, hp, hw), (cl, cp, cw) = out["host"], out["cuda"]
    np.testing.assert_allclose(cl, hl, rtol=1e-2)
    assert np.linalg.norm(cp - hp) <= 1e-2 * np.linalg.norm(hp)
    for (hwi, hbi), (cwi, cbi), (w0, b0) in zip(hw, cw, p0):
        # compare the UPDATES (what the step computed), norm-wise rel 1e-2
        assert np.linalg.norm((cwi - w0) - (hwi - w0)) <= 1e-2 * np.linalg.norm(hwi - w0) + 1e-7
        assert np.linalg.norm((cbi - b0) - (hbi - b0)) <= 1e-2 * np.linalg.norm(hbi - b0) + 1e-7


def test_mlp_step_full_size_properties(cuda_dev, math_mode):
    sizes, batch = (784, 4096, 4096, 10), 8192
    rng = np.random.default_rng(52)
    params = [(CudaTensor.ofNumpy(w), CudaTensor.ofNumpy(b)) for w, b in init_params(rng, sizes)]
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
    losses = []
    for _ in range(4):
        loss, pred = train_step(x, t, params, 0.002)
        losses.append(float(loss.Value))
        pn = pred.toNumpy()
        assert np.isfinite(pn).all() and np.allclose(pn.sum(axis=1), 1.0, atol=1e-4)
        recomputed = float(-(tn * np.log(np.maximum(pn, 1e-45))).sum(axis=1).mean())
        assert abs(recomputed - losses[-1]) <= 1e-3 * abs(recomputed) + 1e-5
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses


def test_mlp_step_fused_equals_unfused(cuda_dev):
    """The FusedElemwise variant of the training step must produce the SAME bits as the operator-by-operator step
    on the device (every fused instruction rounds like the operator it replaces; the GEMMs are the same calls)."""
    sizes, batch = (64, 96, 80, 10), 256
    rng = np.random.default_rng(53)
    p0 = init_params(rng, sizes)
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    outs = []
    for fused in (False, True):
        params = [(CudaTensor.ofNumpy(w.copy()), CudaTensor.ofNumpy(b.copy())) for w, b in p0]
        x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
        for _ in range(3):
            loss, pred = train_step(x, t, params, 0.05, fused=fused)
        outs.append((float(loss.Value), pred.toNumpy(), [(w.toNumpy(), b.toNumpy()) for w, b in params]))
    (l0, p0_, w0), (l1, p1_, w1) = outs
    assert l0 == l1 and np.array_equal(p0_, p1_)
    for (wa, ba), (wb, bb) in zip(w0, w1):
        assert np.array_equal(wa, wb) and np.array_equal(ba, bb)
