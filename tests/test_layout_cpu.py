"""Host-side view logic (deepnet_b200/layout.py = Tensor/Tensor/TensorLayout.fs) checked against numpy's own view
semantics, with no backend involved: every transformation is applied to a layout over `arange(n)` storage and the
elements it addresses must be the elements numpy's equivalent view shows. Seeded sequences of random
transformations (hypothesis-style, but deterministic) cover compositions."""
import numpy as np
import pytest

from deepnet_b200 import layout as TL
from deepnet_b200.layout import TensorLayout


def materialise(lay: TensorLayout, storage: np.ndarray) -> np.ndarray:
    out = np.empty(lay.Shape, dtype=storage.dtype)
    for idx in np.ndindex(*lay.Shape):
        out[idx] = storage[TL.addr(idx, lay)]
    return out


def test_single_transformations_match_numpy():
    base = np.arange(2 * 3 * 4 * 5).reshape(2, 3, 4, 5)
    st = base.reshape(-1)
    lay = TL.newC(base.shape)
    assert TL.isC(lay) and not TL.isF(lay) and TL.hasContiguousMemory(lay)
    assert np.array_equal(materialise(lay, st), base)
    assert np.array_equal(materialise(TL.transpose(lay), st), np.swapaxes(base, 2, 3))
    assert np.array_equal(materialise(TL.swapDim(0, 3, lay), st), np.swapaxes(base, 0, 3))
    # permut[i] is the NEW position of axis i (TensorLayout.fs:358-365) = numpy's transpose with the inverse permutation
    assert np.array_equal(materialise(TL.permuteAxes([2, 0, 3, 1], lay), st), base.transpose(np.argsort([2, 0, 3, 1])))
    assert TL.permuteAxes([2, 0, 3, 1], lay).Shape == (3, 5, 2, 4)
    assert np.array_equal(materialise(TL.reverseAxis(1, lay), st), base[:, ::-1])
    assert np.array_equal(materialise(TL.view([slice(1, None), 2, slice(None), slice(1, 4)], lay), st), base[1:, 2, :, 1:4])
    assert np.array_equal(materialise(TL.view([Ellipsis, 0], lay), st), base[..., 0])
    assert np.array_equal(materialise(TL.view([None, 1, Ellipsis], lay), st), base[None, 1, ...])
    assert np.array_equal(materialise(TL.view([slice(-1, None), slice(0, -1)], lay), st), base[-1:, 0:-1])
    assert np.array_equal(materialise(TL.padLeft(lay), st), base[None])
    assert np.array_equal(materialise(TL.padRight(lay), st), base[..., None])
    assert np.array_equal(materialise(TL.insertAxis(2, lay), st), np.expand_dims(base, 2))
    sq = TL.newC((3, 4, 4))
    sst = np.arange(48)
    assert np.array_equal(materialise(TL.diagAxis(1, 2, sq), sst), np.diagonal(sst.reshape(3, 4, 4), axis1=1, axis2=2))
    assert np.array_equal(materialise(TL.newF((3, 4)), np.arange(12)), np.arange(12).reshape(4, 3).T)
    # broadcasting
    row = TL.newC((1, 5))
    b = TL.broadcastToShape((4, 3, 5), row)
    assert b.Shape == (4, 3, 5) and TL.isBroadcasted(b)
    assert np.array_equal(materialise(b, np.arange(5)), np.broadcast_to(np.arange(5), (4, 3, 5)))
    x, y = TL.broadcastToSameMany([TL.newC((3, 1)), TL.newC((4,))])
    assert x.Shape == y.Shape == (3, 4)
    with pytest.raises(RuntimeError):             # invalidOp -> InvalidOperationException (TensorLayout.fs:214-216)
        TL.broadcastToSameMany([TL.newC((3, 2)), TL.newC((4,))])
    with pytest.raises((RuntimeError, ValueError)):
        TL.broadcastToShape((2,), TL.newC((3,)))
    with pytest.raises(IndexError):
        TL.view([5], lay)
    with pytest.raises(IndexError):
        TL.view([slice(0, 3, 2)], lay)          # Deep.Net ranges have no step (TensorRng.fs:30-38)
    with pytest.raises(ValueError):
        TL.diagAxis(0, 1, lay)                  # 2 != 3


def test_reshape_views_and_copies():
    base = np.arange(24).reshape(2, 3, 4)
    st = base.reshape(-1)
    lay = TL.newC(base.shape)
    for shp in [(6, 4), (2, 12), (24,), (2, 3, 2, 2), (1, 24, 1), (TL.Remainder, 4), (2, TL.Remainder)]:
        r = TL.tryReshape(shp, lay)
        want_shape = tuple(24 // 4 if s == TL.Remainder and shp[1:] == (4,) else s for s in shp)
        assert r is not None and np.array_equal(materialise(r, st), base.reshape(r.Shape))
        assert len(r.Shape) == len(want_shape)
    t = TL.transpose(lay)                       # [2, 4, 3], not C-contiguous
    assert TL.tryReshape((8, 3), t) is None     # would need a copy
    keep = TL.tryReshape((2, 1, 4, 3), t)       # inserting a size-1 dim never needs one
    assert keep is not None and np.array_equal(materialise(keep, st), np.swapaxes(base, 1, 2).reshape(2, 1, 4, 3))
    sl = TL.view([slice(None), slice(0, 2)], lay)  # [2, 2, 4] with pitch 12: dims 1,2 cannot merge with dim 0
    assert TL.tryReshape((4, 4), sl) is None
    with pytest.raises(ValueError):
        TL.tryReshape((5, 5), lay)
    with pytest.raises(ValueError):
        TL.tryReshape((TL.Remainder, TL.Remainder), lay)


def test_random_compositions_match_numpy():
    rng = np.random.default_rng(2024)
    for trial in range(200):
        nd = int(rng.integers(1, 5))
        shape = tuple(int(s) for s in rng.integers(1, 5, size=nd))
        arr = np.arange(int(np.prod(shape))).reshape(shape)
        lay, view = TL.newC(shape), arr
        for _ in range(int(rng.integers(1, 6))):
            op = int(rng.integers(0, 6))
            nd = len(lay.Shape)
            if op == 0 and nd >= 2:
                a1, a2 = (int(x) for x in rng.choice(nd, size=2, replace=False))
                lay, view = TL.swapDim(a1, a2, lay), np.swapaxes(view, a1, a2)
            elif op == 1 and nd >= 1:
                ax = int(rng.integers(0, nd))
                lay, view = TL.reverseAxis(ax, lay), np.flip(view, ax)
            elif op == 2 and nd >= 1:
                ax = int(rng.integers(0, nd))
                n = lay.Shape[ax]
                lo = int(rng.integers(0, n))
                hi = int(rng.integers(lo + 1, n + 1))
                rngs = [slice(None)] * nd
                rngs[ax] = slice(lo, hi)
                lay, view = TL.view(rngs, lay), view[tuple(rngs)]
            elif op == 3 and nd >= 1:
                perm = [int(x) for x in rng.permutation(nd)]
                lay, view = TL.permuteAxes(perm, lay), view.transpose(np.argsort(perm))
            elif op == 4 and nd < 6:
                ax = int(rng.integers(0, nd + 1))
                lay, view = TL.insertAxis(ax, lay), np.expand_dims(view, ax)
            elif op == 5 and nd >= 1:
                ax = int(rng.integers(0, nd))
                if lay.Shape[ax] == 1:
                    target = list(lay.Shape)
                    target[ax] = 3
                    lay, view = TL.broadcastToShape(target, lay), np.broadcast_to(view, target)
        assert lay.Shape == view.shape, trial
        assert np.array_equal(materialise(lay, arr.reshape(-1)), view), trial
        assert lay.NElems == view.size
