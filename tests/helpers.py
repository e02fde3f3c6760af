"""Shared helpers for the parity tests: seeded inputs in the reference benchmark's distribution
(uniform [-50, 50), ints = rounded, bools = u >= 0.5 — Tensor.Benchmark/Benchmark.fs:105-107),
a catalogue of view transformations applied identically on both devices, and the comparison rules from
BASELINE.json north_star (bit-exact for integer/bool/index results, rel 1e-5 for f32 element-wise, ...)."""
from __future__ import annotations

import math

import numpy as np

from deepnet_b200 import CudaTensor, Tensor, dtypes
from oracle.host_tensor import HostTensor

ALL_DTYPES = [dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I8, dtypes.DN_U8, dtypes.DN_I16, dtypes.DN_U16,
              dtypes.DN_I32, dtypes.DN_U32, dtypes.DN_I64, dtypes.DN_U64, dtypes.DN_BOOL]
NUMERIC = [d for d in ALL_DTYPES if d != dtypes.DN_BOOL]
FLOATS = [dtypes.DN_F32, dtypes.DN_F64]
INTS = [d for d in NUMERIC if d not in FLOATS]
SIGNED_INTS = [dtypes.DN_I8, dtypes.DN_I16, dtypes.DN_I32, dtypes.DN_I64]
MAIN_DTYPES = [dtypes.DN_F32, dtypes.DN_F64, dtypes.DN_I32, dtypes.DN_I64]


def rand_array(rng: np.random.Generator, shape, dtype: int, lo=-50.0, hi=50.0) -> np.ndarray:
    npdt = dtypes.to_numpy(dtype)
    u = rng.uniform(lo, hi, size=shape)
    if dtype == dtypes.DN_BOOL:
        return rng.uniform(0, 1, size=shape) >= 0.5
    if dtype in FLOATS:
        return u.astype(npdt)
    r = np.rint(u)
    if np.dtype(npdt).kind == "u":
        r = np.abs(r)
    return r.astype(npdt)


def pair(arr: np.ndarray):
    """The same data as a HostTensor (oracle) and as a CudaTensor."""
    return HostTensor.ofNumpy(arr), CudaTensor.ofNumpy(arr)


# name -> (base shape, view function applied to a Tensor on either device). All results have the same shape per
# entry so that binary ops can mix any two views of one group.
def view_catalogue():
    cat = {
        "contig_1d_tail": ((100003,), lambda t: t),
        "contig_1d_offset1": ((4099,), lambda t: t[1:]),
        "contig_2d": ((67, 128), lambda t: t),
        "odd_2d": ((67, 129), lambda t: t),
        "transposed": ((129, 67), lambda t: t.T),
        "sliced_pitch": ((70, 140), lambda t: t[1:68, 3:132]),
        "reverse_rows": ((67, 129), lambda t: t.reverseAxis(0)),
        "reverse_cols": ((67, 129), lambda t: t.reverseAxis(1)),
        "permuted_3d": ((33, 5, 7), lambda t: t.permuteAxes([2, 0, 1])),
        "contig_3d": ((5, 7, 33), lambda t: t),
        "rank0": ((), lambda t: t),
        "empty": ((0, 5), lambda t: t),
        "single": ((1, 1), lambda t: t),
        "diag": ((40, 40), lambda t: t.diag()),
        "inner_slice": ((50, 64, 3), lambda t: t[:, :, 1]),
    }
    return cat


# groups of views with equal result shape, for binary/ternary operators
SHAPE_GROUPS = {
    (67, 129): ["odd_2d", "transposed", "sliced_pitch", "reverse_rows", "reverse_cols"],
    (5, 7, 33): ["contig_3d", "permuted_3d"],
}


def assert_same(host: Tensor, cuda: Tensor, dtype: int, rtol: float = 0.0, what: str = ""):
    h, c = host.toNumpy(), cuda.toNumpy()
    assert h.shape == c.shape, f"{what}: shape {c.shape} != {h.shape}"
    assert h.dtype == c.dtype, f"{what}: dtype {c.dtype} != {h.dtype}"
    if rtol == 0.0 or dtype not in FLOATS:
        if dtype in FLOATS:
            # bit-exact including NaN payload-insensitivity: NaN == NaN, and +0 / -0 must match exactly
            ok = (h == c) | (np.isnan(h) & np.isnan(c))
            ok &= (np.signbit(h) == np.signbit(c)) | np.isnan(h)
        else:
            ok = h == c
        assert ok.all(), f"{what}: {np.count_nonzero(~ok)} of {ok.size} elements differ; first at " \
                         f"{np.argwhere(~ok)[:1].tolist()} host={h[~ok][:3]} cuda={c[~ok][:3]}"
    else:
        both_nan = np.isnan(h) & np.isnan(c)
        same_inf = np.isinf(h) & (h == c)
        with np.errstate(invalid="ignore", over="ignore"):
            err = np.abs(h.astype(np.float64) - c.astype(np.float64))
            tol = rtol * np.abs(h.astype(np.float64))
        tiny = np.finfo(h.dtype).tiny
        ok = both_nan | same_inf | (err <= tol + tiny)
        assert ok.all(), f"{what}: {np.count_nonzero(~ok)} of {ok.size} elements outside rel {rtol}; " \
                         f"host={h[~ok][:3]} cuda={c[~ok][:3]}"


def reduction_rtol(n: int) -> float:
    """north_star: rel 1e-4 * log2(n) for float32 reductions."""
    return 1e-4 * max(1.0, math.log2(max(2, n)))
