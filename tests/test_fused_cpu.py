"""CPU checks of the fused-expression path (no GPU): the tracer's programs (register allocation, constants, common
subexpressions, limits) and the oracle's dno_fused_elemwise == the oracle's operator-by-operator evaluation."""
import numpy as np
import pytest

from deepnet_b200 import Tensor, dtypes
from deepnet_b200.fused import trace
from deepnet_b200.native import DN_FUSED_BINARY, DN_FUSED_CONST, DN_FUSED_REGS, DN_FUSED_UNARY
from oracle.host_tensor import HostTensor


def run_program(prog, srcs):
    """Reference interpreter in numpy float64 (structure only: registers, ordering)."""
    from deepnet_b200.backend import BINARY_OPS, UNARY_OPS
    r = [None] * DN_FUSED_REGS
    for i, s in enumerate(srcs):
        r[i] = s
    z = None
    un = {"UnaryPlus": lambda x: x, "UnaryMinus": np.negative, "Abs": np.abs, "Sin": np.sin, "Exp": np.exp, "Tanh": np.tanh,
          "Sqrt": np.sqrt, "Log": np.log}
    bi = {"Add": np.add, "Subtract": np.subtract, "Multiply": np.multiply, "Divide": np.divide}
    for kind, op, dst, a, b, imm in prog:
        if kind == DN_FUSED_CONST:
            z = np.full_like(srcs[0], imm)
        elif kind == DN_FUSED_UNARY:
            assert r[a] is not None
            z = un[UNARY_OPS[op]](r[a])
        else:
            assert r[a] is not None and r[b] is not None
            z = bi[BINARY_OPS[op]](r[a], r[b])
        assert 0 <= dst < DN_FUSED_REGS
        r[dst] = z
    return z


def test_tracer_programs():
    x = np.linspace(0.5, 2.0, 7)
    y = np.linspace(-1.0, 1.0, 7)
    cases = [
        (lambda a, b: a * b + a.sin(), (x, y), x * y + np.sin(x)),
        (lambda a, b: a * (1.0 - b * b), (x, y), x * (1 - y * y)),
        (lambda a: a, (x,), x),
        (lambda a: 3.0 - a, (x,), 3 - x),
        (lambda a, b: (a + b) * (a + b) - (a - b) * (a - b), (x, y), 4 * x * y),
        (lambda a, b, c: ((a - b) * c / (abs(c) + 2.5)).tanh(), (x, y, x), np.tanh((x - y) * x / (np.abs(x) + 2.5))),
    ]
    for fn, srcs, want in cases:
        prog = trace(fn, len(srcs))
        np.testing.assert_allclose(run_program(prog, list(srcs)), want, rtol=1e-12, atol=1e-12)
    shared = trace(lambda a, b: (lambda s: s * s)(a + b), 2)        # common subexpression evaluated once
    assert sum(1 for ins in shared if ins[0] == DN_FUSED_BINARY) == 2
    step = 0.25
    p1 = trace(lambda w, g: w - g * step, 2)
    step = 0.5
    p2 = trace(lambda w, g: w - g * step, 2)                         # captured constants are part of the cache key
    assert [i[5] for i in p1 if i[0] == DN_FUSED_CONST] == [0.25] and [i[5] for i in p2 if i[0] == DN_FUSED_CONST] == [0.5]
    with pytest.raises(ValueError):
        trace(lambda a: sum(((a + float(k)) for k in range(20)), a), 1)   # more than 12 instructions


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64])
def test_oracle_fused_is_the_operator_sequence(dtype):
    rng = np.random.default_rng(61)
    npdt = dtypes.to_numpy(dtype)
    a = HostTensor.ofNumpy(rng.uniform(-5, 5, size=(33, 70)).astype(npdt))
    b = HostTensor.ofNumpy(rng.uniform(-5, 5, size=(33, 70)).astype(npdt))
    row = HostTensor.ofNumpy(rng.uniform(-5, 5, size=(1, 70)).astype(npdt))
    for fn, srcs in [(lambda x, y: x * y + x.sin(), (a, b)), (lambda x, y: (x - y).exp() / (abs(y) + 1.0), (a.T, b.T)),
                     (lambda x, y, z: x * (1.0 - y * y) + z, (a, b, row)), (lambda x: -(x * x).sqrt().log(), (a[1:, 3:],))]:
        fused = Tensor.fused(fn, *srcs).toNumpy()
        unfused = fn(*srcs).toNumpy()
        assert np.array_equal(fused, unfused, equal_nan=True)
    a.FillFused(lambda x, y: x - y * 0.125, a, b)
    assert np.isfinite(a.toNumpy()).all()
    with pytest.raises(ValueError):
        a.Backend.FusedElemwise(a, [a], [(DN_FUSED_BINARY, 0, 1, 0, 4, 0.0)])
