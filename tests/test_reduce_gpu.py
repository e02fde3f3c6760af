"""Reduction parity: Sum/Product/Min/Max/All/Any/CountTrue, ArgMin/ArgMax, Find along every axis of several
layouts, against the HostTensor oracle (host semantics: ScalarOps.fs:606-665). Integer / bool / index results are
bit-exact; float Min/Max are exact in value (NaN handling included); float Sum/Product are within
rel 1e-4*log2(n) (north_star), with an absolute floor of the same factor times the mean |element| so that rows
whose sum cancels to ~0 are judged on the scale of their terms."""
import math

import numpy as np
import pytest

from deepnet_b200 import CudaTensor, NotFound, Tensor, dtypes
from helpers import FLOATS, INTS, MAIN_DTYPES, NUMERIC, assert_same, pair, rand_array, reduction_rtol

pytestmark = pytest.mark.gpu

SHAPES_AXES = [
    ((257, 1000), 1), ((257, 1000), 0), ((64, 4096), 1), ((3, 70001), 1), ((70001, 3), 0), ((100003,), 0),
    ((33, 17, 65), 0), ((33, 17, 65), 1), ((33, 17, 65), 2), ((5, 0), 1), ((0, 5), 1), ((7, 1), 1), ((2000, 7), 1),
    ((4, 300000), 1), ((300000, 4), 0),
]


def assert_fold_close(h: Tensor, c: Tensor, src: np.ndarray, axis: int, dtype: int, what: str):
    hn, cn = h.toNumpy(), c.toNumpy()
    assert hn.shape == cn.shape
    if dtype not in FLOATS:
        np.testing.assert_array_equal(cn, hn, err_msg=what)
        return
    n = max(1, src.shape[axis])
    rtol = reduction_rtol(n) if dtype == dtypes.DN_F32 else 1e-12 * max(1.0, math.log2(max(2, n)))
    scale = np.abs(src.astype(np.float64)).mean(axis=axis) if src.size else np.zeros(hn.shape)
    err = np.abs(hn.astype(np.float64) - cn.astype(np.float64))
    tol = rtol * np.abs(hn.astype(np.float64)) + rtol * scale
    assert (err <= tol).all(), f"{what}: max err {err.max()} (tol {tol[err > tol][:3]})"


@pytest.mark.parametrize("shape,axis", SHAPES_AXES)
@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_sum_min_max_arg(cuda_dev, shape, axis, dtype):
    rng = np.random.default_rng(21)
    arr = rand_array(rng, shape, dtype)
    h, c = pair(arr)
    assert_fold_close(h.sumAxis(axis), c.sumAxis(axis), arr, axis, dtype, "sumAxis")
    assert_same(h.maxAxis(axis), c.maxAxis(axis), dtype, what="maxAxis")
    assert_same(h.minAxis(axis), c.minAxis(axis), dtype, what="minAxis")
    assert_same(h.argMaxAxis(axis), c.argMaxAxis(axis), dtypes.DN_I64, what="argMaxAxis")
    assert_same(h.argMinAxis(axis), c.argMinAxis(axis), dtypes.DN_I64, what="argMinAxis")


@pytest.mark.parametrize("dtype", NUMERIC)
def test_all_numeric_dtypes(cuda_dev, dtype):
    rng = np.random.default_rng(22)
    arr = rand_array(rng, (130, 517), dtype, -3, 3)
    h, c = pair(arr)
    for axis in (0, 1):
        assert_fold_close(h.sumAxis(axis), c.sumAxis(axis), arr, axis, dtype, "sumAxis")
        assert_same(h.maxAxis(axis), c.maxAxis(axis), dtype, what="maxAxis")
        assert_same(h.minAxis(axis), c.minAxis(axis), dtype, what="minAxis")
        assert_same(h.argMaxAxis(axis), c.argMaxAxis(axis), dtypes.DN_I64, what="argMaxAxis")
        assert_same(h.argMinAxis(axis), c.argMinAxis(axis), dtypes.DN_I64, what="argMinAxis")
        assert_same(h.findAxis(2, axis), c.findAxis(2, axis), dtypes.DN_I64, what="findAxis")
    if dtype in INTS:  # integer products wrap and are exact in any order
        assert_same(h.productAxis(1), c.productAxis(1), dtype, what="productAxis int")


SUBWORD = [dtypes.DN_I8, dtypes.DN_U8, dtypes.DN_I16, dtypes.DN_U16]


@pytest.mark.parametrize("dtype", SUBWORD)
def test_subword_rows_full_range(cuda_dev, dtype):
    """1- and 2-byte integers take their own row fold (whole 16-byte vectors kept packed — Sum through the
    dot-product unit — plus one element per lane for the misaligned head and the ragged tail): full-range values so
    that the wrapping sums matter, row lengths around every vector / warp-round boundary, rows that start at every
    byte alignment (odd row length and a sliced first column), and a long single row (several parts per row)."""
    info = np.iinfo(dtypes.to_numpy(dtype))
    rng = np.random.default_rng(29)
    for rows, L in [(37, 1), (37, 15), (37, 16), (37, 17), (37, 31), (64, 511), (64, 512), (64, 513), (29, 1000),
                    (11, 2047), (11, 2049), (5, 4099), (3, 70001), (1, 300007)]:
        arr = rng.integers(info.min, int(info.max) + 1, size=(rows, L + 3), dtype=np.int64).astype(info.dtype)
        h0, c0 = pair(arr)
        for h, c, tag in [(h0, c0, "whole rows"), (h0[:, 3:], c0[:, 3:], "first column 3"), (h0[:, 1:L], c0[:, 1:L], "inner slice")]:
            if h.Shape[1] == 0:
                continue
            what = f"{tag} {rows}x{L}"
            assert_same(h.sumAxis(1), c.sumAxis(1), dtype, what="sumAxis " + what)
            assert_same(h.productAxis(1), c.productAxis(1), dtype, what="productAxis " + what)
            assert_same(h.maxAxis(1), c.maxAxis(1), dtype, what="maxAxis " + what)
            assert_same(h.minAxis(1), c.minAxis(1), dtype, what="minAxis " + what)
            assert_same(h.argMaxAxis(1), c.argMaxAxis(1), dtypes.DN_I64, what="argMaxAxis " + what)
            assert_same(h.argMinAxis(1), c.argMinAxis(1), dtypes.DN_I64, what="argMinAxis " + what)
            v = arr[0, L // 2].item()
            assert_same(h.findAxis(v, 1), c.findAxis(v, 1), dtypes.DN_I64, what="findAxis " + what)
    # extremes: rows of the initial value (ArgMax of all-lowest is NotFound), ties (first occurrence)
    arr = np.full((4, 1000), info.min, dtype=info.dtype)
    arr[1, 700] = arr[1, 900] = info.max
    arr[2, :] = info.max
    arr[3, 999] = info.min + 1
    h, c = pair(arr)
    assert_same(h.argMaxAxis(1), c.argMaxAxis(1), dtypes.DN_I64, what="argMax extremes")
    assert_same(h.argMinAxis(1), c.argMinAxis(1), dtypes.DN_I64, what="argMin extremes")
    assert_same(h.maxAxis(1), c.maxAxis(1), dtype, what="max extremes")
    assert_same(h.sumAxis(1), c.sumAxis(1), dtype, what="sum extremes")


@pytest.mark.parametrize("dtype", FLOATS)
def test_product_float(cuda_dev, dtype):
    rng = np.random.default_rng(23)
    arr = rand_array(rng, (64, 3000), dtype, 0.9, 1.1)
    h, c = pair(arr)
    for axis in (0, 1):
        hn, cn = h.productAxis(axis).toNumpy(), c.productAxis(axis).toNumpy()
        rtol = reduction_rtol(arr.shape[axis]) if dtype == dtypes.DN_F32 else 1e-11
        np.testing.assert_allclose(cn, hn, rtol=rtol)


def _special_rows(dtype, L):
    npdt = dtypes.to_numpy(dtype)
    rng = np.random.default_rng(24)
    base = rng.uniform(-50, 50, size=(24, L)).astype(npdt)
    lo = np.finfo(npdt).min
    base[0, :] = np.nan
    base[1, :] = -np.inf
    base[2, :] = lo
    base[3, :] = 7.0
    base[4, 5] = np.nan
    base[5, L - 1] = np.nan
    base[6, 0] = np.nan
    base[7, [3, L // 2, L - 2]] = np.nan
    base[8, :] = 0.0
    base[8, 1::2] = -0.0
    base[9, 10] = base[9].max()
    base[9, 20] = base[9].max()
    base[10, :] = np.inf
    base[11, L // 2:] = np.nan
    base[12, : L // 2] = np.nan
    base[13, L - 33: L - 1] = np.nan
    base[14, [L // 3, L // 3 + 1]] = [np.nan, -np.inf]
    base[15, :] = np.inf
    base[15, L // 2] = np.nan
    base[16, :] = -np.inf
    base[16, 3] = np.nan
    return base


@pytest.mark.parametrize("dtype", FLOATS)
@pytest.mark.parametrize("L", [40, 1000, 5000, 70000])
def test_float_minmax_arg_special_values(cuda_dev, dtype, L):
    """SURVEY.md §8c rules 4-5: NaN resets Max/Min, ArgMax skips NaN, all -inf / all MinValue -> NotFound."""
    arr = _special_rows(dtype, L)
    h, c = pair(arr)
    for view_h, view_c, axis in [(h, c, 1), (h.T.Copy(), c.T.Copy(), 0), (h.T.Copy().T, c.T.Copy().T, 1)]:
        for fn in ("maxAxis", "minAxis"):
            hn, cn = getattr(view_h, fn)(axis).toNumpy(), getattr(view_c, fn)(axis).toNumpy()
            ok = (hn == cn) | (np.isnan(hn) & np.isnan(cn))
            assert ok.all(), f"{fn} rows {np.argwhere(~ok).ravel().tolist()} host={hn[~ok]} cuda={cn[~ok]}"
        assert_same(view_h.argMaxAxis(axis), view_c.argMaxAxis(axis), dtypes.DN_I64, what="argMaxAxis")
        assert_same(view_h.argMinAxis(axis), view_c.argMinAxis(axis), dtypes.DN_I64, what="argMinAxis")
        assert_same(view_h.findAxis(7.0, axis), view_c.findAxis(7.0, axis), dtypes.DN_I64, what="findAxis")
    got = c.argMaxAxis(1).toNumpy()
    assert got[0] == NotFound and got[1] == NotFound and got[2] == NotFound and got[3] == 0


def test_int_arg_not_found(cuda_dev):
    arr = np.full((3, 100), np.iinfo(np.int32).min, dtype=np.int32)
    arr[1, 40] = -5
    arr[2, :] = np.iinfo(np.int32).max
    h, c = pair(arr)
    assert_same(h.argMaxAxis(1), c.argMaxAxis(1), dtypes.DN_I64, what="argMax int")
    assert_same(h.argMinAxis(1), c.argMinAxis(1), dtypes.DN_I64, what="argMin int")
    assert c.argMaxAxis(1).toNumpy().tolist() == [NotFound, 40, 0]
    assert c.argMinAxis(1).toNumpy().tolist() == [0, 0, NotFound]


@pytest.mark.parametrize("shape,axis", [((257, 1000), 1), ((257, 1000), 0), ((100003,), 0), ((3, 70001), 1),
                                        ((33, 17, 65), 1), ((6, 0), 1)])
def test_bool_folds(cuda_dev, shape, axis):
    rng = np.random.default_rng(25)
    arr = rng.uniform(0, 1, size=shape) >= 0.02
    if arr.size:
        arr.reshape(-1)[:: max(1, arr.size // 7)] = True
    h, c = pair(arr)
    assert_same(h.allAxis(axis), c.allAxis(axis), dtypes.DN_BOOL, what="allAxis")
    assert_same((~h).anyAxis(axis), (~c).anyAxis(axis), dtypes.DN_BOOL, what="anyAxis")
    assert_same(h.countTrueAxis(axis), c.countTrueAxis(axis), dtypes.DN_I64, what="countTrueAxis")
    assert_same(h.findAxis(False, axis), c.findAxis(False, axis), dtypes.DN_I64, what="findAxis bool")


@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_whole_tensor_and_views(cuda_dev, dtype):
    rng = np.random.default_rng(26)
    arr = rand_array(rng, (300, 401), dtype)
    h, c = pair(arr)
    for vh, vc in [(h, c), (h.T, c.T), (h[3:200, 5:], c[3:200, 5:]), (h.reverseAxis(1), c.reverseAxis(1))]:
        if dtype in FLOATS:
            hs, cs = float(vh.sum()), float(vc.sum())
            scale = float(np.abs(arr).mean())
            rt = reduction_rtol(vh.NElems) if dtype == dtypes.DN_F32 else 1e-11
            assert abs(hs - cs) <= rt * abs(hs) + rt * scale
        else:
            assert vh.sum() == vc.sum()
        assert vh.max() == vc.max() and vh.min() == vc.min()
        assert vh.argMax() == vc.argMax() and vh.argMin() == vc.argMin()
        for axis in (0, 1):
            assert_same(vh.maxAxis(axis), vc.maxAxis(axis), dtype, what="maxAxis view")
            assert_same(vh.argMinAxis(axis), vc.argMinAxis(axis), dtypes.DN_I64, what="argMinAxis view")
    v = arr[17, 33]
    assert h.tryFind(v) == c.tryFind(v)
    assert c.tryFind(12345) is None


def test_broadcast_source_and_strided_target(cuda_dev, host_dev):
    rng = np.random.default_rng(27)
    arr = rand_array(rng, (1, 700), dtypes.DN_I32)
    h, c = pair(arr)
    hb, cb = h.broadcastTo((50, 700)), c.broadcastTo((50, 700))
    assert_same(hb.sumAxis(1), cb.sumAxis(1), dtypes.DN_I32, what="sum of broadcast rows")
    assert_same(hb.sumAxis(0), cb.sumAxis(0), dtypes.DN_I32, what="sum over broadcast axis")
    ht = Tensor.zeros((700, 3), dtypes.DN_I32, host_dev)
    ct = Tensor.zeros((700, 3), dtypes.DN_I32, cuda_dev)
    ht[:, 1].FillSumAxis(0, hb)
    ct[:, 1].FillSumAxis(0, cb)
    assert_same(ht, ct, dtypes.DN_I32, what="FillSumAxis into a strided target")
