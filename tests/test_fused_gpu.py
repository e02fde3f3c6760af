"""dn_fused_elemwise: a traced expression evaluated in one pass must equal the operator-by-operator
evaluation — bit for bit against the same device's unfused calls for arithmetic programs, and against the HostTensor
oracle under the element-wise tolerances of north_star (bit-exact for IEEE-exact operators, rel 1e-5 / 1e-12 when a
transcendental function is involved)."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes
from helpers import pair, rand_array

pytestmark = pytest.mark.gpu

def mx(a, b):
    return a.maxElemwise(b) if not isinstance(a, Tensor) else Tensor.maxElemwise(a, b)


def mn(a, b):
    return a.minElemwise(b) if not isinstance(a, Tensor) else Tensor.minElemwise(a, b)


EXACT = {
    "a*b+a": (lambda a, b: a * b + a, 2),
    "tanh'": (lambda dh, h: dh * (1.0 - h * h), 2),
    "sgd": (lambda w, g: w - g * 0.01, 2),
    "three": (lambda a, b, c: (a - b) * c / (abs(c) + 2.5), 3),
    "minmax": (lambda a, b: mx(a, b) - mn(a, b * 0.5), 2),
    "neg-sqrt-abs": (lambda a: -(abs(a).sqrt()) + a.floor() - a.ceil() + a.round() * a.truncate(), 1),
    "mod": (lambda a, b: a % (abs(b) + 1.0), 2),
    "reuse": (lambda a, b: (a + b) * (a + b) - (a - b) * (a - b), 2),
}
TRANSCENDENTAL = {
    "a*b+sin(a)": (lambda a, b: a * b + a.sin(), 2),
    "softmax-num": (lambda z, c: (z - c).exp(), 2),
    "log-cosh": (lambda a: (a * 0.1).cosh().log() + (a * 0.02).tanh(), 1),
    "pow": (lambda a, b: (abs(a) + 1.0) ** (b * 0.05), 2),
}


def unfused(fn, *ts):
    """The same Python expression evaluated with ordinary Tensor operators (one backend call each)."""
    return fn(*ts)


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
@pytest.mark.parametrize("name", list(EXACT) + list(TRANSCENDENTAL))
def test_fused_matches_unfused_and_oracle(cuda_dev, dtype, name):
    fn, nsrc = (EXACT | TRANSCENDENTAL)[name]
    rng = np.random.default_rng(51)
    arrs = [rand_array(rng, (67, 256), dtype) for _ in range(nsrc)]
    hs, cs = zip(*[pair(a) for a in arrs])
    got = Tensor.fused(fn, *cs).toNumpy()
    want_dev = unfused(fn, *cs).toNumpy()
    want_host = Tensor.fused(fn, *hs).toNumpy()
    if name in EXACT:
        assert np.array_equal(got, want_dev, equal_nan=True), "fused != unfused on the device"
        assert np.array_equal(got, want_host, equal_nan=True), "fused != oracle"
    else:
        rtol = 1e-5 if dtype == dtypes.DN_F32 else 1e-12
        np.testing.assert_allclose(got, want_dev, rtol=rtol * 0.2, atol=0)
        np.testing.assert_allclose(got, want_host, rtol=rtol, atol=rtol)
    # the oracle's fused result is the oracle's unfused result, always
    assert np.array_equal(want_host, unfused(fn, *hs).toNumpy(), equal_nan=True)


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_fused_views_broadcast_in_place(cuda_dev, dtype):
    rng = np.random.default_rng(52)
    a, b = rand_array(rng, (70, 140), dtype), rand_array(rng, (70, 140), dtype)
    row, col = rand_array(rng, (1, 140), dtype), rand_array(rng, (70, 1), dtype)
    (ha, ca), (hb, cb), (hr, cr), (hc, cc) = pair(a), pair(b), pair(row), pair(col)
    fn = lambda x, y, z: (x - y) * z + x
    for what, (H, C) in {
        "broadcast row/col": ((ha, hr, hc), (ca, cr, cc)),
        "transposed": ((ha.T, hb.T, ha.T), (ca.T, cb.T, ca.T)),
        "sliced + reversed": ((ha[1:, 3:], hb[1:, 3:].reverseAxis(1), hb[1:, 3:]), (ca[1:, 3:], cb[1:, 3:].reverseAxis(1), cb[1:, 3:])),
        "scalar operand": ((ha, Tensor.scalar(2.5, dtype, ha.Dev), hb), (ca, Tensor.scalar(2.5, dtype, ca.Dev), cb)),
    }.items():
        assert np.array_equal(Tensor.fused(fn, *C).toNumpy(), Tensor.fused(fn, *H).toNumpy()), what
    # in place: the target is also a source (f3.FillMultiply f3 e idiom, Tensor.Sample/Program.fs:186)
    ha.FillFused(lambda x, y: x * y - y, ha, hb)
    ca.FillFused(lambda x, y: x * y - y, ca, cb)
    assert np.array_equal(ca.toNumpy(), ha.toNumpy())
    # tails and tiny tensors
    for n in (1, 5, 31, 1000003):
        v = rand_array(rng, (n,), dtype)
        hv, cv = pair(v)
        assert np.array_equal(Tensor.fused(lambda x: x * x + 1.0, cv).toNumpy(), Tensor.fused(lambda x: x * x + 1.0, hv).toNumpy())


def test_fused_program_validation(cuda_dev):
    from deepnet_b200.native import DN_FUSED_BINARY, DN_FUSED_UNARY, NotSupportedException
    a = CudaTensor.zeros((8,), dtypes.DN_F32)
    with pytest.raises(ValueError):   # reads register 3 before anything wrote it
        a.Backend.FusedElemwise(a, [a], [(DN_FUSED_BINARY, 0, 1, 0, 3, 0.0)])
    with pytest.raises(ValueError):   # destination register out of range
        a.Backend.FusedElemwise(a, [a], [(DN_FUSED_UNARY, 1, 9, 0, 0, 0.0)])
    i = CudaTensor.zeros((8,), dtypes.DN_I32)
    with pytest.raises(NotSupportedException):
        i.Backend.FusedElemwise(i, [i], [(DN_FUSED_UNARY, 1, 1, 0, 0, 0.0)])
    with pytest.raises(ValueError):   # more live values than registers
        Tensor.fused(lambda x: ((((x + 1.0) * (x + 2.0)) * ((x + 3.0) * (x + 4.0))) * (((x + 5.0) * (x + 6.0)) * ((x + 7.0) * (x + 8.0)))) *
                     ((((x + 9.0) * (x + 10.0)) * ((x + 11.0) * (x + 12.0))) * (((x + 13.0) * (x + 14.0)) * ((x + 15.0) * (x + 16.0)))), a)


def test_interpreter_stays_covered():
    """Programs that have a precompiled form (a*b + sin a, the MLP's element-wise chains) no longer reach the
    interpreter; the DN_FUSED_INTERPRET=1 test hook (read once per process, hence the child process) sends everything
    through it again, so both evaluators are held to the same bit-identity checks of this file and of the MLP step."""
    import os
    import subprocess
    import sys
    if os.environ.get("DN_FUSED_INTERPRET") == "1":
        pytest.skip("already running under the hook")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, DN_FUSED_INTERPRET="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), os.path.join(here, "test_mlp_gpu.py"),
                          "-k", "fused", "-q", "-m", "gpu", "-x"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
