"""Indexing parity: Gather, Scatter, MaskedGet/MaskedSet, TrueIndices, countTrue against the HostTensor oracle
(host semantics: ScalarOps.fs:583-604,667-707). All integer / bool / index results are bit-exact; float Scatter sums
duplicates in a different order than the host's sequential loop and is compared at rel 1e-5."""
import numpy as np
import pytest

from deepnet_b200 import CudaTensor, NoMask, Tensor, dtypes
from helpers import ALL_DTYPES, FLOATS, MAIN_DTYPES, NUMERIC, assert_same, pair, rand_array

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_gather_1d(cuda_dev, dtype):
    rng = np.random.default_rng(31)
    hs, cs = pair(rand_array(rng, (5003,), dtype))
    hi, ci = pair(rng.integers(0, 5003, size=(20011,), dtype=np.int64))
    assert_same(Tensor.gather([hi], hs), Tensor.gather([ci], cs), dtype, what="gather 1-D")


@pytest.mark.parametrize("dtype", MAIN_DTYPES)
def test_gather_nd_with_none_and_views(cuda_dev, dtype):
    rng = np.random.default_rng(32)
    hs, cs = pair(rand_array(rng, (37, 41, 5), dtype))
    # all three dims specified, index tensors of shape [50, 30]
    idx = [rng.integers(0, n, size=(50, 30), dtype=np.int64) for n in (37, 41, 5)]
    hi, ci = zip(*[pair(i) for i in idx])
    assert_same(Tensor.gather(list(hi), hs), Tensor.gather(list(ci), cs), dtype, what="gather 3 idx")
    # None on dim 0 and 1 (target dims 0,1 map through), index on dim 2; target shape [37, 41]
    h2, c2 = pair(rng.integers(0, 5, size=(37, 41), dtype=np.int64))
    assert_same(Tensor.gather([None, None, h2], hs), Tensor.gather([None, None, c2], cs), dtype, what="gather None")
    # source is a transposed / reversed view, index tensors broadcast from [50, 1] and [1, 30]
    hsv, csv = hs.permuteAxes([2, 0, 1]).reverseAxis(0), cs.permuteAxes([2, 0, 1]).reverseAxis(0)  # [41, 5, 37]
    ia, ib, ic = (rng.integers(0, 41, size=(50, 1), dtype=np.int64), rng.integers(0, 5, size=(1, 30), dtype=np.int64),
                  rng.integers(0, 37, size=(50, 30), dtype=np.int64))
    (ha, ca), (hb, cb), (hc, cc) = pair(ia), pair(ib), pair(ic)
    assert_same(Tensor.gather([ha, hb, hc], hsv), Tensor.gather([ca, cb, cc], csv), dtype, what="gather views")
    # gather into a strided target
    ht, ct = pair(np.zeros((50, 60), dtype=dtypes.to_numpy(dtype)))
    ht[:, 10:40].FillGather([ha.broadcastTo((50, 30)), hb.broadcastTo((50, 30)), hc], hsv)
    ct[:, 10:40].FillGather([ca.broadcastTo((50, 30)), cb.broadcastTo((50, 30)), cc], csv)
    assert_same(ht, ct, dtype, what="FillGather strided target")


def test_gather_out_of_range(cuda_dev):
    """CudaKernels.fs:335-342: with Cfg.Stacktrace the op synchronises and raises IndexOutOfRangeException."""
    cs = CudaTensor.ofNumpy(np.arange(10, dtype=np.float32))
    ci = CudaTensor.ofNumpy(np.array([1, 10, 3], dtype=np.int64))
    cuda_dev.SetStacktrace(True)
    try:
        with pytest.raises(IndexError):
            Tensor.gather([ci], cs)
        with pytest.raises(IndexError):
            Tensor.scatter([CudaTensor.ofNumpy(np.array([-1, 2, 3], dtype=np.int64))], (10,),
                           CudaTensor.ofNumpy(np.ones(3, dtype=np.float32)))
        # and the flag is cleared again
        assert Tensor.gather([CudaTensor.ofNumpy(np.array([1, 2], dtype=np.int64))], cs).toNumpy().tolist() == [1, 2]
    finally:
        cuda_dev.SetStacktrace(False)


@pytest.mark.parametrize("dtype", NUMERIC)
def test_scatter_1d_collisions(cuda_dev, dtype):
    rng = np.random.default_rng(33)
    src = rand_array(rng, (20011,), dtype, -3, 3)
    hs, cs = pair(src)
    for idx in (rng.integers(0, 503, size=(20011,), dtype=np.int64),            # heavy collisions
                rng.permutation(20011).astype(np.int64),                         # collision-free
                np.minimum(rng.geometric(0.05, size=20011), 600).astype(np.int64)):  # hot spots
        hi, ci = pair(idx)
        h, c = Tensor.scatter([hi], (20011,), hs), Tensor.scatter([ci], (20011,), cs)
        if dtype in FLOATS:
            # atomics add duplicates in a different order than the host's sequential loop: compare on the scale
            # of the terms that were summed into each cell (rel 1e-5 for f32, north_star)
            scale = np.zeros(20011, dtype=np.float64)
            np.add.at(scale, idx, np.abs(src.astype(np.float64)))
            err = np.abs(h.toNumpy().astype(np.float64) - c.toNumpy().astype(np.float64))
            assert (err <= (1e-5 if dtype == dtypes.DN_F32 else 1e-13) * scale).all(), "scatter 1-D float"
        else:
            assert_same(h, c, dtype, what="scatter 1-D")


@pytest.mark.parametrize("dtype", [dtypes.DN_I64, dtypes.DN_I32, dtypes.DN_F64])
def test_scatter_nd(cuda_dev, dtype):
    """The doc example (Tensor.fs:2168-2185) generalised: 2-D source, both target dims indexed, plus None."""
    rng = np.random.default_rng(34)
    hs, cs = pair(rand_array(rng, (30, 40), dtype, -3, 3))
    (h0, c0), (h1, c1) = pair(rng.integers(0, 17, size=(30, 40), dtype=np.int64)), \
        pair(rng.integers(0, 19, size=(30, 40), dtype=np.int64))
    rt = 1e-11 if dtype == dtypes.DN_F64 else 0.0
    assert_same(Tensor.scatter([h0, h1], (17, 19), hs), Tensor.scatter([c0, c1], (17, 19), cs), dtype, rt, "scatter 2-D")
    assert_same(Tensor.scatter([None, h1], (30, 19), hs), Tensor.scatter([None, c1], (30, 19), cs), dtype, rt,
                "scatter None")
    assert_same(Tensor.scatter([h0.T, None], (17, 30), hs.T), Tensor.scatter([c0.T, None], (17, 30), cs.T), dtype, rt,
                "scatter transposed")


@pytest.mark.parametrize("shape", [(100003,), (67, 129), (129, 67), (5, 7, 33), (4096,), (1,), (0,), (3, 0)])
@pytest.mark.parametrize("p", [0.5, 0.01, 1.0, 0.0])
def test_true_idx_and_count(cuda_dev, shape, p):
    rng = np.random.default_rng(35)
    arr = rng.uniform(0, 1, size=shape) < p
    h, c = pair(arr)
    assert h.countTrue() == c.countTrue() == int(arr.sum())
    assert_same(h.trueIdx(), c.trueIdx(), dtypes.DN_I64, what="trueIdx")
    if len(shape) == 2 and arr.size:
        assert h.T.countTrue() == c.T.countTrue()
        assert_same(h.T.trueIdx(), c.T.trueIdx(), dtypes.DN_I64, what="trueIdx transposed")
        assert_same(h[1:, 2:].trueIdx(), c[1:, 2:].trueIdx(), dtypes.DN_I64, what="trueIdx sliced")


@pytest.mark.parametrize("shape", [(1024, 1031), (2048, 1024), (70000, 16), (33, 40000)])
@pytest.mark.parametrize("p", [0.5, 0.03, 1.0])
def test_true_idx_large_matrix(cuda_dev, shape, p):
    """>= 2^20 positions (the 32768-positions-per-CTA configuration of the compaction kernel), incl. rows that are
    not a multiple of 16 long, strided masks and a target with fewer rows than there are true elements."""
    rng = np.random.default_rng(39)
    arr = rng.uniform(0, 1, size=shape) < p
    h, c = pair(arr)
    want = h.trueIdx()
    assert_same(want, c.trueIdx(), dtypes.DN_I64, what="trueIdx large")
    assert_same(h.T.trueIdx(), c.T.trueIdx(), dtypes.DN_I64, what="trueIdx large transposed")
    assert_same(h[1:, 3:].trueIdx(), c[1:, 3:].trueIdx(), dtypes.DN_I64, what="trueIdx large sliced")
    # capped target: only the first rows are written, the rest of the buffer keeps its sentinel
    n = want.Shape[0]
    if n > 1000:
        cap = n - 777
        buf = CudaTensor.ofNumpy(np.full((n, 2), -7, dtype=np.int64))
        part = buf[0:cap]
        part.Backend.TrueIndices(part, c)
        got = buf.toNumpy()
        assert (got[:cap] == want.toNumpy()[:cap]).all() and (got[cap:] == -7).all()


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_masked_get_set_whole_tensor(cuda_dev, dtype):
    rng = np.random.default_rng(36)
    for shape in [(100003,), (67, 129)]:
        arr = rand_array(rng, shape, dtype)
        mask = rng.uniform(0, 1, size=shape) < 0.5
        (h, c), (hm, cm) = pair(arr), pair(mask)
        hg, cg = h.M(hm), c.M(cm)
        assert hg.Shape == cg.Shape == (int(mask.sum()),)
        assert_same(hg, cg, dtype, what="MaskedGet")
        vals = rand_array(rng, (int(mask.sum()),), dtype)
        hv, cv = pair(vals)
        h.SetM([hm], hv)
        c.SetM([cm], cv)
        assert_same(h, c, dtype, what="MaskedSet")
        hs, cs = pair(rand_array(rng, (), dtype))
        h.SetM([hm], hs)   # scalar value broadcast to every selected element
        c.SetM([cm], cs)
        assert_same(h, c, dtype, what="MaskedSet broadcast value")


@pytest.mark.parametrize("dtype", [dtypes.DN_F32, dtypes.DN_I64])
def test_masked_per_dimension(cuda_dev, dtype):
    """Tensor.fs:3120-3146: one mask per dimension selects the cartesian product; NoMask keeps a dimension."""
    rng = np.random.default_rng(37)
    arr = rand_array(rng, (23, 31, 7), dtype)
    m0, m1, m2 = (rng.uniform(0, 1, size=n) < 0.6 for n in (23, 31, 7))
    (h, c), (h0, c0), (h1, c1), (h2, c2) = pair(arr), pair(m0), pair(m1), pair(m2)
    for hm, cm in [([h0, h1, h2], [c0, c1, c2]), ([h0, NoMask, h2], [c0, NoMask, c2]),
                   ([NoMask, NoMask, h2], [NoMask, NoMask, c2]), ([h0, NoMask, NoMask], [c0, NoMask, NoMask])]:
        hg, cg = h.M(*hm), c.M(*cm)
        assert_same(hg, cg, dtype, what="MaskedGet per-dim")
        hv, cv = pair(rand_array(rng, hg.Shape, dtype))
        h.SetM(hm, hv)
        c.SetM(cm, cv)
        assert_same(h, c, dtype, what="MaskedSet per-dim")
    # a 2-D mask covering the first two dims of a 3-D tensor (frontend flattens it: Tensor.fs:3047-3050)
    m01 = rng.uniform(0, 1, size=(23, 31)) < 0.3
    h01, c01 = pair(m01)
    assert_same(h.M(h01, NoMask), c.M(c01, NoMask), dtype, what="MaskedGet 2-D mask + NoMask")
    # masked get on a non-contiguous view (reshape copies first: Tensor.fs:466-480)
    hT, cT = h.permuteAxes([2, 0, 1]), c.permuteAxes([2, 0, 1])
    mT = rng.uniform(0, 1, size=hT.Shape) < 0.5
    hmT, cmT = pair(mT)
    assert_same(hT.M(hmT), cT.M(cmT), dtype, what="MaskedGet permuted view")


def test_masked_per_dimension_surplus_positions_stay_untouched(cuda_dev):
    """A dense side LARGER than the number of selected elements along a masked dimension (the backend call allows
    it; the frontend sizes it exactly): the host walks the selected elements only (ScalarOps.fs:667-707), so the
    surplus positions of a MaskedGet target keep their contents and the surplus values of a MaskedSet are not
    stored anywhere — in particular not at index 0 (round-1 review)."""
    rng = np.random.default_rng(41)
    arr = rng.integers(1, 1000, size=(9, 12), dtype=np.int64)
    m0 = np.array([0, 1, 1, 0, 0, 1, 0, 0, 0], dtype=bool)           # 3 rows selected
    m1 = rng.uniform(0, 1, size=12) < 0.5
    n1 = int(m1.sum())
    src, c0, c1 = CudaTensor.ofNumpy(arr), CudaTensor.ofNumpy(m0), CudaTensor.ofNumpy(m1)
    # MaskedGet into a target with 5 rows (2 surplus) and n1 + 1 columns (1 surplus)
    trg = CudaTensor.ofNumpy(np.full((5, n1 + 1), -7, dtype=np.int64))
    src.Backend.MaskedGet(trg, src, [c0, c1])
    want = np.full((5, n1 + 1), -7, dtype=np.int64)
    want[:3, :n1] = arr[m0][:, m1]
    np.testing.assert_array_equal(trg.toNumpy(), want)
    # MaskedSet from values with surplus rows / columns: only the selected cells change
    vals = rng.integers(-1000, -1, size=(5, n1 + 1), dtype=np.int64)
    full = CudaTensor.ofNumpy(arr.copy())
    full.Backend.MaskedSet(full, [c0, c1], CudaTensor.ofNumpy(vals))
    want = arr.copy()
    want[np.ix_(m0, m1)] = vals[:3, :n1]
    np.testing.assert_array_equal(full.toNumpy(), want)


def test_compaction_with_32768_position_tiles(cuda_dev):
    """Inputs of >= ~58 M positions use 32768 positions per CTA (4 staging rounds). DN_COMPACT_ROUNDS=4 forces that
    configuration so the compaction tests above also cover it at sizes the oracle finishes in seconds."""
    import os
    import subprocess
    import sys
    if os.environ.get("DN_COMPACT_ROUNDS"):
        pytest.skip("already running under a forced configuration")
    env = dict(os.environ, DN_COMPACT_ROUNDS="4")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                          "true_idx or masked"], env=env, capture_output=True, text=True, timeout=900,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def test_scatter_binned_path_on_the_small_cases():
    """The bin-then-accumulate Scatter (large dense targets) is forced on for every Scatter case of this file through
    the DN_SCATTER_BINNED=1 test hook (read once per process, hence the child process); test_configs_gpu.py covers it
    at the sizes it is meant for."""
    import os
    import subprocess
    import sys
    if os.environ.get("DN_SCATTER_BINNED") == "1":
        pytest.skip("already running under the hook")
    env = dict(os.environ, DN_SCATTER_BINNED="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-k", "scatter", "-q", "-m", "gpu", "-x"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
