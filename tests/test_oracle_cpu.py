"""A second, independent restatement of the HostTensor rules (rules 1-9) in numpy / plain Python,
checked against the C++ oracle on seeded inputs and views. The documented known answers (tests/test_golden.py) pin
the oracle on the reference's own examples; this file pins the RULES on thousands of elements, so that a slip in
host_oracle.cpp cannot hide behind the small doc examples."""
import math

import numpy as np
import pytest

from deepnet_b200 import NotFound, Tensor, dtypes
from oracle.host_tensor import HostTensor

RNG = np.random.default_rng(2718)


def H(a):
    return HostTensor.ofNumpy(np.asarray(a))


def test_rule1_float32_functions_are_evaluated_in_double():
    x = RNG.uniform(-3, 3, size=(37, 53)).astype(np.float32)
    pos = np.abs(x) + np.float32(0.01)
    unit = (x / np.float32(3.01)).astype(np.float32)
    for name, fn, arg in [("sin", np.sin, x), ("cos", np.cos, x), ("tan", np.tan, x), ("exp", np.exp, x),
                          ("log", np.log, pos), ("log10", np.log10, pos), ("asin", np.arcsin, unit),
                          ("acos", np.arccos, unit), ("atan", np.arctan, x), ("sinh", np.sinh, x), ("cosh", np.cosh, x),
                          ("tanh", np.tanh, x), ("sqrt", np.sqrt, pos)]:
        got = getattr(H(arg), name)().toNumpy()
        want = fn(arg.astype(np.float64)).astype(np.float32)     # (float)fn((double)x), ScalarPrimitives.fs:72-155
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, want, rtol=2e-7, atol=0, err_msg=name)   # libm vs numpy: <= 1 ulp of double
    p = (H(pos) ** H(unit)).toNumpy()
    np.testing.assert_allclose(p, np.power(pos.astype(np.float64), unit.astype(np.float64)).astype(np.float32), rtol=2e-7)


def test_rules_2_3_rounding_and_sign():
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 2.4999, 1e10, np.nan, np.inf, -np.inf, -0.0, 0.0])
    assert np.array_equal(H(v).round().toNumpy(), np.rint(v), equal_nan=True)          # half to even
    assert np.array_equal(H(v).ceil().toNumpy(), np.ceil(v), equal_nan=True)
    assert np.array_equal(H(v).floor().toNumpy(), np.floor(v), equal_nan=True)
    assert np.array_equal(H(v).truncate().toNumpy(), np.trunc(v), equal_nan=True)
    sgn = H(v).sgn().toNumpy()
    want = np.where(np.isnan(v), 0.0, np.sign(v))                                        # Sgn(NaN) = 0 on the host
    assert np.array_equal(sgn, want)
    i = np.array([-7, 0, 9, np.iinfo(np.int32).min], dtype=np.int32)
    assert np.array_equal(H(i).sgn().toNumpy(), np.sign(i))


def test_rule6_integer_arithmetic_wraps_and_truncates():
    for npdt in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint32, np.uint64):
        info = np.iinfo(npdt)
        a = RNG.integers(info.min, info.max, size=(29, 31), dtype=npdt, endpoint=True)
        b = RNG.integers(info.min, info.max, size=(29, 31), dtype=npdt, endpoint=True)
        with np.errstate(over="ignore"):
            assert np.array_equal((H(a) + H(b)).toNumpy(), a + b)
            assert np.array_equal((H(a) - H(b)).toNumpy(), a - b)
            assert np.array_equal((H(a) * H(b)).toNumpy(), a * b)
        small = RNG.integers(1, 50, size=(29, 31)).astype(npdt)
        if info.min < 0:
            small = np.where(RNG.uniform(size=small.shape) < 0.5, -small, small).astype(npdt)
            a2 = np.where(a == info.min, info.min + 1, a).astype(npdt)           # MinValue / -1 throws on the host
        else:
            a2 = a
        q = (H(a2) / H(small)).toNumpy()
        r = (H(a2) % H(small)).toNumpy()
        wq = np.trunc(a2.astype(np.float64) / small.astype(np.float64)) if info.bits < 64 else None
        if wq is not None:
            assert np.array_equal(q, wq.astype(npdt))                              # truncation toward zero
        assert np.array_equal(q.astype(object) * small.astype(object) + r.astype(object), a2.astype(object))
        assert (np.abs(r.astype(object)) < np.abs(small.astype(object))).all()
        if info.min < 0:
            assert ((r == 0) | (np.sign(r) == np.sign(a2))).all()                  # remainder has the dividend's sign
            m = np.array([info.min, -1, 0, 5], dtype=npdt)
            with np.errstate(over="ignore"):
                assert np.array_equal(abs(H(m)).toNumpy(), np.abs(m))              # wraps at MinValue (Vector.Abs)
                assert np.array_equal((-H(m)).toNumpy(), -m)


def fold_minmax(row, is_max, init):
    res = init
    for v in row:                      # ScalarOps.fs:620-628: if res > v then res else v
        res = res if ((res > v) if is_max else (res < v)) else v
    return res


def test_rules_4_5_8_reductions_follow_the_sequential_fold():
    f = RNG.uniform(-50, 50, size=(41, 23)).astype(np.float32)
    f[3, 5] = np.nan
    f[4, 22] = np.nan
    f[5, :] = np.nan
    f[6, :] = -np.inf
    f[7, :] = 12.5
    f[8, 2] = f[8, 9] = 99.0
    h = H(f)
    fmax, fmin = np.finfo(np.float32).max, np.finfo(np.float32).min
    mx, mn = h.maxAxis(1).toNumpy(), h.minAxis(1).toNumpy()
    for r in range(f.shape[0]):
        want = fold_minmax(f[r], True, np.float32(fmin))
        assert (np.isnan(want) and np.isnan(mx[r])) or want == mx[r], r
        want = fold_minmax(f[r], False, np.float32(fmax))
        assert (np.isnan(want) and np.isnan(mn[r])) or want == mn[r], r
    am, an = h.argMaxAxis(1).toNumpy(), h.argMinAxis(1).toNumpy()
    for r in range(f.shape[0]):
        best, idx = np.float32(fmin), NotFound                  # strict >, first occurrence, NaNs never win
        for k, v in enumerate(f[r]):
            if v > best:
                best, idx = v, k
        assert am[r] == idx, r
        best, idx = np.float32(fmax), NotFound
        for k, v in enumerate(f[r]):
            if v < best:
                best, idx = v, k
        assert an[r] == idx, r
    assert am[6] == NotFound and am[5] == NotFound and am[8] == 2
    # Sum / Product fold left to right IN the element type (ScalarOps.fs:610-618)
    g = RNG.uniform(-2, 2, size=(9, 200)).astype(np.float32)
    s, p = H(g).sumAxis(1).toNumpy(), H(g[:, :20]).productAxis(1).toNumpy()
    for r in range(g.shape[0]):
        acc = np.float32(0)
        for v in g[r]:
            acc = np.float32(acc + v)
        assert s[r] == acc
        acc = np.float32(1)
        for v in g[r, :20]:
            acc = np.float32(acc * v)
        assert p[r] == acc
    i = RNG.integers(-2 ** 62, 2 ** 62, size=(7, 50))
    with np.errstate(over="ignore"):
        assert np.array_equal(H(i).sumAxis(1).toNumpy(), i.sum(axis=1))              # wraps like .NET unchecked
    b = RNG.uniform(size=(13, 17)) < 0.8
    assert np.array_equal(H(b).allAxis(1).toNumpy(), b.all(axis=1)) and np.array_equal(H(b).anyAxis(0).toNumpy(), b.any(axis=0))
    assert np.array_equal(H(b).countTrueAxis(1).toNumpy(), b.sum(axis=1))
    assert np.array_equal(H(i % 5).findAxis(3, 1).toNumpy(),
                          np.array([next((k for k, v in enumerate(row) if v == 3), NotFound) for row in (i % 5)]))


def test_rule9_indexing_walks_in_logical_row_major_order():
    src = RNG.integers(-1000, 1000, size=(6, 7, 5))
    view = np.swapaxes(src, 0, 2)[::-1]                            # a transposed + reversed VIEW, shape [5, 7, 6]
    hv = H(src).swapDim(0, 2).reverseAxis(0)
    mask = RNG.uniform(size=view.shape) < 0.4
    hm = H(mask)
    assert np.array_equal(hv.M(hm).toNumpy(), view[mask])                              # numpy boolean indexing is row-major
    assert np.array_equal(hm.trueIdx().toNumpy(), np.argwhere(mask))
    assert hm.countTrue() == int(mask.sum())
    i0 = RNG.integers(0, 5, size=(4, 9))
    i1 = RNG.integers(0, 7, size=(4, 9))
    i2 = RNG.integers(0, 6, size=(4, 9))
    assert np.array_equal(Tensor.gather([H(i0), H(i1), H(i2)], hv).toNumpy(), view[i0, i1, i2])
    vals = RNG.integers(-9, 9, size=(4, 9))
    want = np.zeros((5, 7, 6), dtype=np.int64)
    np.add.at(want, (i0, i1, i2), vals)                                                # duplicates are summed
    assert np.array_equal(Tensor.scatter([H(i0), H(i1), H(i2)], (5, 7, 6), H(vals)).toNumpy(), want)
    tgt = np.array(view)
    new = RNG.integers(-5, 5, size=int(mask.sum()))
    ht = H(tgt)
    ht.SetM([hm], H(new))
    tgt[mask] = new
    assert np.array_equal(ht.toNumpy(), tgt)


def test_rule7_convert_truncates_and_fill_incrementing():
    f = np.array([-2.9, -2.5, -0.5, 0.5, 2.5, 2.9, 1e9], dtype=np.float64)
    assert np.array_equal(H(f).convert(dtypes.DN_I32).toNumpy(), f.astype(np.int32))            # toward zero
    assert np.array_equal(H(f[:6]).convert(dtypes.DN_I8).toNumpy(), f[:6].astype(np.int32).astype(np.int8))
    i = RNG.integers(-2 ** 40, 2 ** 40, size=33)
    assert np.array_equal(H(i).convert(dtypes.DN_F32).toNumpy(), i.astype(np.float32))
    assert np.array_equal(H(i).convert(dtypes.DN_I16).toNumpy(), i.astype(np.int16))           # unchecked narrowing
    assert np.array_equal(H(np.array([0.0, -0.0, 3.5, np.nan])).convert(dtypes.DN_BOOL).toNumpy(), np.array([False, False, True, True]))
    t = Tensor.arange(HostTensor.Dev, 1.0, 0.25, 3.0, dtypes.DN_F32)
    k = np.arange(t.Shape[0], dtype=np.float32)
    assert np.array_equal(t.toNumpy(), np.float32(1.0) + np.float32(0.25) * k)                    # start + incr * conv(pos)
    assert math.isclose(float(t.toNumpy()[-1]), 2.75)
