"""Seeded randomised parity: random ranks, shapes, view transformations (permute, slice, reverse, broadcast,
new axes) and dtypes; every element-wise family, reductions along a random axis and indexing ops are compared
with the HostTensor oracle. Integer / bool / index results bit-exact, floats per north_star tolerances."""
import numpy as np
import pytest

from deepnet_b200 import NoMask, Tensor, dtypes
from helpers import FLOATS, assert_same, pair, rand_array, reduction_rtol

pytestmark = pytest.mark.gpu

DTYPES = [dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I32, dtypes.DN_I64, dtypes.DN_I16, dtypes.DN_U8]


def random_view(rng, h, c):
    """Applies the same random chain of view ops to the host and cuda tensors."""
    for _ in range(rng.integers(0, 4)):
        nd = h.NDims
        if nd == 0:
            break
        op = rng.integers(0, 4)
        if op == 0 and nd >= 2:
            perm = list(rng.permutation(nd))
            h, c = h.permuteAxes(perm), c.permuteAxes(perm)
        elif op == 1:
            ax = int(rng.integers(0, nd))
            h, c = h.reverseAxis(ax), c.reverseAxis(ax)
        elif op == 2:
            rngs = []
            for s in h.Shape:
                if s >= 2 and rng.random() < 0.6:
                    a = int(rng.integers(0, s // 2))
                    b = int(rng.integers(a + 1, s + 1))
                    rngs.append(slice(a, b))
                else:
                    rngs.append(slice(None))
            h, c = h[tuple(rngs)], c[tuple(rngs)]
        elif op == 3 and nd >= 2 and rng.random() < 0.3:
            ax = int(rng.integers(0, nd))
            sl = tuple(slice(None) if d != ax else 0 for d in range(nd))
            h, c = h[sl], c[sl]
    return h, c


def random_shape(rng):
    nd = int(rng.integers(1, 5))
    return tuple(int(rng.choice([1, 2, 3, 5, 8, 17, 33, 64, 100])) for _ in range(nd))


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_elementwise_and_reduce(cuda_dev, seed):
    rng = np.random.default_rng(1000 + seed)
    dtype = DTYPES[seed % len(DTYPES)]
    shape = random_shape(rng)
    ha, ca = pair(rand_array(rng, shape, dtype, -9, 9))
    hb, cb = pair(rand_array(rng, shape, dtype, -9, 9))
    # the same view chain on both operands' hosts/cudas keeps shapes compatible: build views from a common chain
    state = rng.bit_generator.state
    hav, cav = random_view(rng, ha, ca)
    rng.bit_generator.state = state
    hbv, cbv = random_view(rng, hb, cb)
    assert hav.Shape == hbv.Shape
    what = f"seed {seed} shape {shape} -> {hav.Shape} strides {hav.Stride}"
    assert_same(hav + hbv, cav + cbv, dtype, what="add " + what)
    assert_same(hav * hbv, cav * cbv, dtype, what="mul " + what)
    assert_same(Tensor.maxElemwise(hav, hbv), Tensor.maxElemwise(cav, cbv), dtype, what="max " + what)
    assert_same(hav.le(hbv), cav.le(cbv), dtypes.DN_BOOL, what="le " + what)
    assert_same(Tensor.ifThenElse(hav.gt(hbv), hav, hbv), Tensor.ifThenElse(cav.gt(cbv), cav, cbv), dtype,
                what="select " + what)
    assert_same(hav.Copy(), cav.Copy(), dtype, what="copy " + what)
    assert_same(hav.convert(dtypes.DN_F64), cav.convert(dtypes.DN_F64), dtypes.DN_F64, what="convert " + what)
    # in-place on a view
    hav.FillSubtract(hav, hbv)
    cav.FillSubtract(cav, cbv)
    assert_same(ha, ca, dtype, what="in-place sub through a view " + what)
    if hav.NDims >= 1:
        ax = int(rng.integers(0, hav.NDims))
        assert_same(hav.maxAxis(ax), cav.maxAxis(ax), dtype, what=f"maxAxis {ax} " + what)
        assert_same(hav.argMinAxis(ax), cav.argMinAxis(ax), dtypes.DN_I64, what=f"argMinAxis {ax} " + what)
        assert_same(hav.findAxis(3, ax), cav.findAxis(3, ax), dtypes.DN_I64, what=f"findAxis {ax} " + what)
        hs, cs = hav.sumAxis(ax), cav.sumAxis(ax)
        if dtype in FLOATS:
            n = max(2, hav.Shape[ax])
            rt = reduction_rtol(n) if dtype == dtypes.DN_F32 else 1e-11
            np.testing.assert_allclose(cs.toNumpy(), hs.toNumpy(), rtol=rt, atol=rt * 9 * np.sqrt(n))
        else:
            assert_same(hs, cs, dtype, what=f"sumAxis {ax} " + what)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_indexing(cuda_dev, seed):
    rng = np.random.default_rng(2000 + seed)
    dtype = [dtypes.DN_I64, dtypes.DN_F32, dtypes.DN_I32][seed % 3]
    shape = random_shape(rng)
    hs, cs = pair(rand_array(rng, shape, dtype, -9, 9))
    hsv, csv = random_view(rng, hs, cs)
    if hsv.NDims == 0 or 0 in hsv.Shape:
        return
    what = f"seed {seed} {shape} -> {hsv.Shape} {hsv.Stride}"
    # gather with one random index tensor per source dim
    tshape = tuple(int(x) for x in rng.integers(1, 20, size=int(rng.integers(1, 3))))
    idx = [rng.integers(0, n, size=tshape, dtype=np.int64) for n in hsv.Shape]
    hi, ci = zip(*[pair(i) for i in idx])
    assert_same(Tensor.gather(list(hi), hsv), Tensor.gather(list(ci), csv), dtype, what="gather " + what)
    # mask ops on the view
    mask = rng.uniform(0, 1, size=hsv.Shape) < 0.4
    hm, cm = pair(mask)
    assert hm.countTrue() == cm.countTrue()
    assert_same(hm.trueIdx(), cm.trueIdx(), dtypes.DN_I64, what="trueIdx " + what)
    assert_same(hsv.M(hm), csv.M(cm), dtype, what="maskedGet " + what)
    if hsv.NDims >= 2:
        m0 = rng.uniform(0, 1, size=hsv.Shape[0]) < 0.5
        h0, c0 = pair(m0)
        masks_h, masks_c = [h0] + [NoMask] * (hsv.NDims - 1), [c0] + [NoMask] * (hsv.NDims - 1)
        assert_same(hsv.M(*masks_h), csv.M(*masks_c), dtype, what="maskedGet dim0 " + what)
    if dtype != dtypes.DN_F32:  # integer scatter is exact in any order
        sidx = [rng.integers(0, 7, size=hsv.Shape, dtype=np.int64) for _ in range(2)]
        (h0i, c0i), (h1i, c1i) = pair(sidx[0]), pair(sidx[1])
        assert_same(Tensor.scatter([h0i, h1i], (7, 7), hsv), Tensor.scatter([c0i, c1i], (7, 7), csv), dtype,
                    what="scatter " + what)
