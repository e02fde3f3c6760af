"""Element-wise parity: every ITensorBackend element-wise member on CudaTensor vs the HostTensor oracle, over a
catalogue of views (contiguous, transposed, broadcast, sliced, reversed, permuted, diagonal, rank 0, empty).
Mirrors the reference's host-vs-CUDA comparisons (Tensor.Test/CudaTests.fs:54-177, which run these ops on both
devices; here every result is asserted). Tolerances: BASELINE.json north_star."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, dtypes
from deepnet_b200.native import NotSupportedException
from helpers import (ALL_DTYPES, FLOATS, INTS, MAIN_DTYPES, NUMERIC, SHAPE_GROUPS, SIGNED_INTS, assert_same, pair,
                     rand_array, view_catalogue)

pytestmark = pytest.mark.gpu

CAT = view_catalogue()
FLOAT_UNARY = ["log", "log10", "exp", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "sqrt",
               "ceil", "floor", "round", "truncate"]
EXACT_FLOAT_UNARY = {"sqrt", "ceil", "floor", "round", "truncate"}


def _domain(fn, rng, shape, dtype):
    if fn in ("log", "log10", "sqrt"):
        return rand_array(rng, shape, dtype, 0.01, 50.0)
    if fn in ("asin", "acos"):
        return rand_array(rng, shape, dtype, -1.0, 1.0)
    if fn in ("exp", "sinh", "cosh"):
        return rand_array(rng, shape, dtype, -20.0, 20.0)
    return rand_array(rng, shape, dtype)


@pytest.mark.parametrize("view", list(CAT))
@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_copy_neg_abs_all_views(cuda_dev, view, dtype):
    rng = np.random.default_rng(1)
    shape, fn = CAT[view]
    h, c = pair(rand_array(rng, shape, dtype))
    hv, cv = fn(h), fn(c)
    assert_same(hv.Copy(), cv.Copy(), dtype, what=f"copy {view}")
    assert_same(-hv, -cv, dtype, what=f"neg {view}")
    assert_same(abs(hv), abs(cv), dtype, what=f"abs {view}")
    assert_same(+hv, +cv, dtype, what=f"plus {view}")


@pytest.mark.parametrize("fn", FLOAT_UNARY)
@pytest.mark.parametrize("dtype", FLOATS)
def test_float_unary(cuda_dev, fn, dtype):
    rng = np.random.default_rng(2)
    for shape, view in [((100003,), lambda t: t), ((129, 67), lambda t: t.T), ((70, 140), lambda t: t[1:68, 3:132])]:
        h, c = pair(_domain(fn, rng, shape, dtype))
        rtol = 0.0 if fn in EXACT_FLOAT_UNARY else (1e-5 if dtype == dtypes.DN_F32 else 1e-12)
        assert_same(getattr(view(h), fn)(), getattr(view(c), fn)(), dtype, rtol, what=fn)


def test_float_unary_special_values(cuda_dev):
    vals = np.array([0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 3.5, 1e30, -1e30, np.inf, -np.inf, np.nan,
                     1e-40, 2.4999998, 8388609.0], dtype=np.float32)
    for dtype, arr in ((dtypes.DN_F32, vals), (dtypes.DN_F64, vals.astype(np.float64))):
        h, c = pair(arr)
        for fn in ["sgn", "round", "ceil", "floor", "truncate", "abs", "sqrt"]:
            assert_same(getattr(h, fn)(), getattr(c, fn)(), dtype, what=fn)
        assert_same(h.isFinite(), c.isFinite(), dtypes.DN_BOOL, what="isFinite")
        assert_same(-h, -c, dtype, what="neg")


@pytest.mark.parametrize("dtype", NUMERIC)
def test_unary_int_and_sgn(cuda_dev, dtype):
    rng = np.random.default_rng(3)
    arr = rand_array(rng, (1000,), dtype)
    info = np.iinfo(arr.dtype) if dtype in INTS else None
    if info is not None:
        arr[:4] = [info.min, info.max, 0, 1]
    h, c = pair(arr)
    assert_same(-h, -c, dtype, what="neg")
    assert_same(abs(h), abs(c), dtype, what="abs")
    if dtype in (dtypes.DN_I16, dtypes.DN_I32, dtypes.DN_I64, dtypes.DN_F32, dtypes.DN_F64):
        assert_same(h.sgn(), c.sgn(), dtype, what="sgn")
    else:
        with pytest.raises(NotSupportedException):
            c.sgn()
    if dtype in INTS:
        with pytest.raises(NotSupportedException):
            c.sin()


@pytest.mark.parametrize("op", ["__add__", "__sub__", "__mul__", "__truediv__", "__mod__", "maxElemwise",
                                "minElemwise"])
@pytest.mark.parametrize("dtype", NUMERIC)
def test_binary_contiguous_all_dtypes(cuda_dev, op, dtype):
    rng = np.random.default_rng(4)
    a = rand_array(rng, (10007,), dtype)
    b = rand_array(rng, (10007,), dtype)
    if op in ("__truediv__", "__mod__") and dtype in INTS:
        b[b == 0] = 3  # x/0 throws on the host: outside the parity domain
    ha, ca = pair(a)
    hb, cb = pair(b)
    if op in ("maxElemwise", "minElemwise"):
        hr, cr = getattr(Tensor, op)(ha, hb), getattr(Tensor, op)(ca, cb)
    else:
        hr, cr = getattr(ha, op)(hb), getattr(ca, op)(cb)
    assert_same(hr, cr, dtype, what=op)  # +,-,*,/,fmod,min,max are IEEE-exact: bit-exact for floats too


@pytest.mark.parametrize("dtype", FLOATS)
def test_power(cuda_dev, dtype):
    rng = np.random.default_rng(5)
    ha, ca = pair(rand_array(rng, (5000,), dtype, 0.01, 8.0))
    hb, cb = pair(rand_array(rng, (5000,), dtype, -4.0, 4.0))
    assert_same(ha ** hb, ca ** cb, dtype, 1e-5 if dtype == dtypes.DN_F32 else 1e-12, what="pow")
    with pytest.raises(NotSupportedException):
        CudaTensor.zeros((3,), dtypes.DN_I32) ** CudaTensor.zeros((3,), dtypes.DN_I32)


@pytest.mark.parametrize("shape", list(SHAPE_GROUPS))
@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_binary_view_combinations(cuda_dev, shape, dtype):
    rng = np.random.default_rng(6)
    names = SHAPE_GROUPS[shape]
    for va in names:
        for vb in names:
            sa, fa = CAT[va]
            sb, fb = CAT[vb]
            ha, ca = pair(rand_array(rng, sa, dtype))
            hb, cb = pair(rand_array(rng, sb, dtype))
            assert_same(fa(ha) + fb(hb), fa(ca) + fb(cb), dtype, what=f"{va}+{vb}")
            assert_same(fa(ha).lt(fb(hb)), fa(ca).lt(fb(cb)), dtypes.DN_BOOL, what=f"{va}<{vb}")


@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_broadcast_row_col_scalar(cuda_dev, dtype):
    rng = np.random.default_rng(7)
    ha, ca = pair(rand_array(rng, (61, 132), dtype))
    hrow, crow = pair(rand_array(rng, (1, 132), dtype))
    hcol, ccol = pair(rand_array(rng, (61, 1), dtype))
    hvec, cvec = pair(rand_array(rng, (132,), dtype))
    assert_same(ha + hrow, ca + crow, dtype, what="a+row")
    assert_same(ha * hcol, ca * ccol, dtype, what="a*col")
    assert_same(hcol - hrow, ccol - crow, dtype, what="col-row (outer)")
    assert_same(ha + hvec, ca + cvec, dtype, what="a+vec (padLeft)")
    assert_same(ha * 3, ca * 3, dtype, what="a*scalar")
    assert_same(7 - ha, 7 - ca, dtype, what="scalar-a")
    assert_same(ha.T + hcol.T, ca.T + ccol.T, dtype, what="a.T + col.T")


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_comparisons(cuda_dev, dtype):
    rng = np.random.default_rng(8)
    a = rand_array(rng, (4099,), dtype, -3, 3)
    b = rand_array(rng, (4099,), dtype, -3, 3)
    if dtype in FLOATS:
        a[:3] = [np.nan, 1.0, np.nan]
        b[:3] = [np.nan, np.nan, 1.0]
    ha, ca = pair(a)
    hb, cb = pair(b)
    for op in ["eq", "ne", "lt", "le", "gt", "ge"]:
        assert_same(getattr(ha, op)(hb), getattr(ca, op)(cb), dtypes.DN_BOOL, what=op)
    assert_same(ha.isFinite(), ca.isFinite(), dtypes.DN_BOOL, what="isFinite")


def test_logic(cuda_dev):
    rng = np.random.default_rng(9)
    for shape, view in [((100003,), lambda t: t), ((129, 67), lambda t: t.T)]:
        ha, ca = pair(rand_array(rng, shape, dtypes.DN_BOOL))
        hb, cb = pair(rand_array(rng, shape, dtypes.DN_BOOL))
        ha, ca, hb2, cb2 = view(ha), view(ca), view(hb).Copy(), view(cb).Copy()
        assert_same(~ha, ~ca, dtypes.DN_BOOL, what="not")
        assert_same(ha & hb2, ca & cb2, dtypes.DN_BOOL, what="and")
        assert_same(ha | hb2, ca | cb2, dtypes.DN_BOOL, what="or")
        assert_same(ha ^ hb2, ca ^ cb2, dtypes.DN_BOOL, what="xor")
    with pytest.raises(NotSupportedException):
        ca + ca


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_if_then_else(cuda_dev, dtype):
    rng = np.random.default_rng(10)
    hc, cc = pair(rand_array(rng, (67, 129), dtypes.DN_BOOL))
    ht, ct = pair(rand_array(rng, (129, 67), dtype))
    hf, cf = pair(rand_array(rng, (1, 129), dtype))
    assert_same(Tensor.ifThenElse(hc, ht.T, hf), Tensor.ifThenElse(cc, ct.T, cf), dtype, what="ifThenElse")


@pytest.mark.parametrize("src", ALL_DTYPES)
@pytest.mark.parametrize("dst", ALL_DTYPES)
def test_convert(cuda_dev, src, dst):
    rng = np.random.default_rng(11)
    lo, hi = (-50.0, 50.0)
    if dst in (dtypes.DN_U8, dtypes.DN_U16, dtypes.DN_U32, dtypes.DN_U64) and src in FLOATS + SIGNED_INTS:
        lo = 0.0  # negative -> unsigned is platform-defined on the host: outside the parity domain
    arr = rand_array(rng, (4099,), src, lo, hi)
    h, c = pair(arr)
    assert_same(h.convert(dst), c.convert(dst), dst, what="convert")
    h2, c2 = pair(rand_array(rng, (33, 65), src, lo, hi))
    assert_same(h2.T.convert(dst), c2.T.convert(dst), dst, what="convert transposed")


@pytest.mark.parametrize("dtype", NUMERIC)
def test_fill(cuda_dev, host_dev, dtype):
    for shape in [(1000,), (17, 33), ()]:
        h, c = Tensor.filled(shape, 7, dtype, host_dev), Tensor.filled(shape, 7, dtype, cuda_dev)
        assert_same(h, c, dtype, what="FillConst")
    h = Tensor.empty((25, 40), dtype, host_dev)
    c = Tensor.empty((25, 40), dtype, cuda_dev)
    h.T[1:, :].FillConst(3)
    c.T[1:, :].FillConst(3)
    h[:, 0:1].FillConst(1)
    c[:, 0:1].FillConst(1)
    assert_same(h, c, dtype, what="FillConst on views")


def test_fill_incrementing(cuda_dev, host_dev):
    for dev_pair in [(host_dev, cuda_dev)]:
        h, c = (Tensor.counting(d, 1000) for d in dev_pair)
        assert_same(h, c, dtypes.DN_I64, what="counting")
        h, c = (Tensor.arange(d, 1.0, 0.1, 2.0) for d in dev_pair)
        assert_same(h, c, dtypes.DN_F64, what="arange")
        h, c = (Tensor.linspace(d, 1.0, 2.0, 5, dtypes.DN_F32) for d in dev_pair)
        assert_same(h, c, dtypes.DN_F32, what="linspace")
        for order in ("C", "F"):
            h, c = (Tensor.empty((13, 7), dtypes.DN_I32, d, order=order) for d in dev_pair)
            h.FillIncrementing(5, 3)
            c.FillIncrementing(5, 3)
            assert_same(h, c, dtypes.DN_I32, what=f"FillIncrementing 2-D {order}")


def test_in_place_and_fill_variants(cuda_dev):
    """Guide-Operations.md:103-117: f3.FillMultiply d e; f3.FillMultiply f3 e."""
    rng = np.random.default_rng(12)
    hd, cd = pair(rand_array(rng, (4097,), dtypes.DN_F64))
    he, ce = pair(rand_array(rng, (4097,), dtypes.DN_F64))
    hf, cf = pair(np.zeros(4097))
    hf.FillMultiply(hd, he)
    cf.FillMultiply(cd, ce)
    assert_same(hf, cf, dtypes.DN_F64, what="FillMultiply")
    hf.FillMultiply(hf, he)
    cf.FillMultiply(cf, ce)
    assert_same(hf, cf, dtypes.DN_F64, what="in-place FillMultiply")


def test_setitem_and_transfer_roundtrip(cuda_dev):
    """CudaTests.fs:29-51: transfer round trip, including a non-row-major layout."""
    rng = np.random.default_rng(13)
    arr = rand_array(rng, (11, 13, 5), dtypes.DN_F32)
    c = CudaTensor.ofNumpy(arr)
    np.testing.assert_array_equal(c.toNumpy(), arr)
    np.testing.assert_array_equal(c.permuteAxes([1, 2, 0]).toNumpy(), arr.transpose(2, 0, 1))
    h, c = pair(arr)
    hv, cv = pair(rand_array(rng, (11, 1, 5), dtypes.DN_F32))
    h[:, 2:9, :] = hv
    c[:, 2:9, :] = cv
    assert_same(h, c, dtypes.DN_F32, what="SetRng")
    assert c.Item(3, 4, 1) == h.Item(3, 4, 1)
    c.SetItem((3, 4, 1), 42.0)
    assert c.Item(3, 4, 1) == 42.0


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I8, dtypes.DN_I16, dtypes.DN_I64, dtypes.DN_BOOL])
def test_peeled_rows_and_reversed_vector_paths(cuda_dev, dtype):
    """Misaligned row starts with aligned pitches run as vector body + scalar head/tail columns; sources whose
    innermost stride is -1 are loaded as vectors and reversed in registers. Every combination must match the
    host walk element for element (rows long enough for the widest pack: 32 bools)."""
    rng = np.random.default_rng(41)
    R, C = 37, 512
    (ha, ca), (hb, cb) = pair(rand_array(rng, (R, C), dtype)), pair(rand_array(rng, (R, C), dtype))
    ht, ct = pair(np.zeros((R, C), dtype=dtypes.to_numpy(dtype)))
    is_bool = dtype == dtypes.DN_BOOL

    def op(t, x, y):
        if is_bool:
            t.FillXor(x, y)
        else:
            t.FillAdd(x, y)

    slices = [(slice(1, None), slice(1, None)), (slice(0, None), slice(3, 500)), (slice(2, 30), slice(5, 509)),
              (slice(0, None), slice(0, 401)), (slice(1, None), slice(16, 500))]
    for rs, cs in slices:
        op(ht[rs, cs], ha[rs, cs], hb[rs, cs])
        op(ct[rs, cs], ca[rs, cs], cb[rs, cs])
        assert_same(ht, ct, dtype, what=f"peeled add {rs} {cs}")
        # different misalignment per operand: falls back to the scalar kernel, same answer
        rows = ht[rs, cs].Shape[0]
        n = ht[rs, cs].Shape[1]
        if n + 2 <= C:
            op(ht[rs, cs], ha[rs, 0:n], hb[rs, 2:n + 2])
            op(ct[rs, cs], ca[rs, 0:n], cb[rs, 2:n + 2])
            assert_same(ht, ct, dtype, what=f"mixed misalignment {rs} {cs}")
        # reversed innermost axis on one or both sources, alone and combined with the peel
        op(ht[rs, cs], ha[rs, cs].reverseAxis(1), hb[rs, cs])
        op(ct[rs, cs], ca[rs, cs].reverseAxis(1), cb[rs, cs])
        assert_same(ht, ct, dtype, what=f"reversed source {rs} {cs}")
        op(ht[rs, cs], ha[rs, cs].reverseAxis(1), hb[rs, cs].reverseAxis(1).reverseAxis(0))
        op(ct[rs, cs], ca[rs, cs].reverseAxis(1), cb[rs, cs].reverseAxis(1).reverseAxis(0))
        assert_same(ht, ct, dtype, what=f"both reversed {rs} {cs}")
        assert rows > 0
    # reversed TARGET (the planner flips it, which reverses the sources instead) and comparison into bool
    op(ht.reverseAxis(1), ha, hb.reverseAxis(1))
    op(ct.reverseAxis(1), ca, cb.reverseAxis(1))
    assert_same(ht, ct, dtype, what="reversed target")
    hm, cm = pair(np.zeros((R, C), dtype=np.bool_))
    hm[1:, 1:].FillLess(ha[1:, 1:].reverseAxis(1), hb[1:, 1:])
    cm[1:, 1:].FillLess(ca[1:, 1:].reverseAxis(1), cb[1:, 1:])
    assert_same(hm, cm, dtypes.DN_BOOL, what="compare reversed + peeled")
    ht[1:, 1:].FillIfThenElse(hm[1:, 1:], ha[1:, 1:].reverseAxis(1), hb[1:, 1:])
    ct[1:, 1:].FillIfThenElse(cm[1:, 1:], ca[1:, 1:].reverseAxis(1), cb[1:, 1:])
    assert_same(ht, ct, dtype, what="select reversed + peeled")
    # 1-D views with an offset (head peel of a flat tensor)
    (h1, c1), (h2, c2) = pair(rand_array(rng, (10007,), dtype)), pair(rand_array(rng, (10007,), dtype))
    ho, co = pair(np.zeros((10007,), dtype=dtypes.to_numpy(dtype)))
    op(ho[3:], h1[3:], h2[3:].reverseAxis(0))
    op(co[3:], c1[3:], c2[3:].reverseAxis(0))
    assert_same(ho, co, dtype, what="1-D offset + reversed")
