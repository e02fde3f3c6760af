"""ctypes binding of the C ABI in include/dn_tensor.h.

`CApi` binds one shared library exporting the operator entry points under a prefix: `dn_` for
libdeepnet_b200.so (the product), `dno_` for the CPU oracle used by the tests. The struct layout below IS the ABI
(`dn_tensor`, include/dn_tensor.h) — the F# binding declares the same sequential struct (fsharp/Native.fs).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import dtypes
from .layout import TensorLayout

DN_MAX_DIMS = 8

# dn_status (include/dn_tensor.h) and the exception the reference raises in the same situation.
DN_OK, DN_ERR_INVALID_ARG, DN_ERR_UNSUPPORTED, DN_ERR_OUT_OF_MEMORY, DN_ERR_INDEX_OUT_OF_RANGE, DN_ERR_CUDA, \
    DN_ERR_NO_DEVICE, DN_ERR_SHAPE_MISMATCH, DN_ERR_SINGULAR_MATRIX = range(9)


class CudaException(RuntimeError):
    """ManagedCuda CudaException — CudaBackend.fs:28-38."""


class OutOfCudaMemoryException(MemoryError):
    """CudaUtils.fs:183-210."""


class SingularMatrixException(ArithmeticError):
    """Tensor.SingularMatrixException (Tensor/Tensor/Tensor.fs): raised by invert for non-invertible matrices."""


class NotSupportedException(NotImplementedError):
    """System.NotSupportedException — CudaBackend.fs:126-129."""


_EXC = {
    DN_ERR_INVALID_ARG: ValueError,
    DN_ERR_UNSUPPORTED: NotSupportedException,
    DN_ERR_OUT_OF_MEMORY: OutOfCudaMemoryException,
    DN_ERR_INDEX_OUT_OF_RANGE: IndexError,
    DN_ERR_CUDA: CudaException,
    DN_ERR_NO_DEVICE: CudaException,
    DN_ERR_SHAPE_MISMATCH: RuntimeError,  # InvalidOperationException
    DN_ERR_SINGULAR_MATRIX: SingularMatrixException,
}


class dn_tensor(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("offset", C.c_int64),
        ("ndims", C.c_int32),
        ("dtype", C.c_int32),
        ("shape", C.c_int64 * DN_MAX_DIMS),
        ("stride", C.c_int64 * DN_MAX_DIMS),
    ]


class dn_fused_instr(C.Structure):
    """include/dn_tensor.h `dn_fused_instr`."""
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32),
                ("imm", C.c_double)]


DN_FUSED_UNARY, DN_FUSED_BINARY, DN_FUSED_CONST = range(3)
DN_FUSED_REGS, DN_FUSED_MAX_INSTRS, DN_FUSED_MAX_SRCS = 6, 12, 3


def make_desc(base: int, layout: TensorLayout, dtype: int) -> dn_tensor:
    if layout.NDims > DN_MAX_DIMS:
        raise NotSupportedException(f"tensors of rank {layout.NDims} > {DN_MAX_DIMS} are not supported")
    d = dn_tensor()
    d.base = base
    d.offset = layout.Offset
    d.ndims = layout.NDims
    d.dtype = dtype
    for i, (s, st) in enumerate(zip(layout.Shape, layout.Stride)):
        d.shape[i] = s
        d.stride[i] = st
    return d


_P = C.POINTER(dn_tensor)
_PP = C.POINTER(_P)

# name -> argtypes, shared by both prefixes (operator entry points only)
_OPERATOR_SIGNATURES = {
    "fill_const": [_P, C.c_void_p],
    "fill_incrementing": [_P, C.c_void_p, C.c_void_p],
    "copy": [_P, _P],
    "convert": [_P, _P],
    "unary": [C.c_int32, _P, _P],
    "binary": [C.c_int32, _P, _P, _P],
    "compare": [C.c_int32, _P, _P, _P],
    "is_finite": [_P, _P],
    "if_then_else": [_P, _P, _P, _P],
    "reduce_last_axis": [C.c_int32, _P, _P],
    "arg_reduce_last_axis": [C.c_int32, _P, _P],
    "find_last_axis": [C.c_void_p, _P, _P],
    "gather": [_P, _PP, C.c_int32, _P],
    "scatter": [_P, _PP, C.c_int32, _P],
    "count_true": [_P, C.POINTER(C.c_int64)],
    "masked_get": [_P, _P, _PP, C.c_int32],
    "masked_set": [_P, _PP, C.c_int32, _P],
    "true_indices": [_P, _P],
    "vec_vec_dot": [_P, _P, _P],
    "mat_vec_dot": [_P, _P, _P],
    "mat_mat_dot": [_P, _P, _P],
    "batched_mat_mat_dot": [_P, _P, _P],
    "batched_invert": [_P, _P],
    "fused_elemwise": [_P, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32],
}

# device / storage entry points exported only by the product library
_DEVICE_SIGNATURES = {
    "init": [C.c_int32],
    "device_count": [C.POINTER(C.c_int32)],
    "set_device": [C.c_int32],
    "get_device": [C.POINTER(C.c_int32)],
    "set_stream": [C.c_void_p],
    "get_stream": [C.POINTER(C.c_void_p)],
    "sync": [],
    "set_check_errors": [C.c_int32],
    "poll_index_error": [C.POINTER(C.c_int32)],
    "alloc": [C.c_int64, C.POINTER(C.c_void_p)],
    "free": [C.c_void_p],
    "set_math_mode": [C.c_int32],
    "get_math_mode": [C.POINTER(C.c_int32)],
    "free_deferred": [C.c_void_p],
    "release_stream": [C.c_void_p],
    "host_register": [C.c_void_p, C.c_int64],
    "host_unregister": [C.c_void_p],
    "transfer_h2d": [_P, _P],
    "transfer_d2h": [_P, _P],
    "alloc_host": [C.c_int64, C.POINTER(C.c_void_p)],
    "free_host": [C.c_void_p],
    "memset_zero": [C.c_void_p, C.c_int64],
    "memcpy_h2d": [C.c_void_p, C.c_void_p, C.c_int64],
    "memcpy_d2h": [C.c_void_p, C.c_void_p, C.c_int64],
    "memcpy_d2d": [C.c_void_p, C.c_void_p, C.c_int64],
    "memcpy_d2h_async": [C.c_void_p, C.c_void_p, C.c_int64],
    "event_create": [C.POINTER(C.c_void_p)],
    "event_destroy": [C.c_void_p],
    "event_record": [C.c_void_p],
    "stream_wait_event": [C.c_void_p],
    "get_item": [_P, C.POINTER(C.c_int64), C.c_void_p],
    "set_item": [_P, C.POINTER(C.c_int64), C.c_void_p],
    "arg_reduce_combine": [C.c_int32, _P, _P, _P],
    # leading-axis sharding (include/dn_tensor.h "Multi-GPU")
    "shard_group_create": [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                           C.POINTER(C.c_void_p)],
    "shard_group_handle": [C.c_void_p, C.c_int32, C.c_void_p],
    "shard_group_connect": [C.c_void_p, C.c_void_p],
    "shard_group_destroy": [C.c_void_p],
    "shard_set_stream": [C.c_void_p, C.c_int32, C.c_void_p],
    "shard_sync": [C.c_void_p, C.c_int32],
    "shard_slab": [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "shard_heap_alloc": [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)],
    "shard_heap_reset": [C.c_void_p, C.c_int32],
    "shard_barrier": [C.c_void_p, C.c_int32],
    "shard_group_start": [C.c_void_p],
    "shard_group_end": [C.c_void_p],
    "shard_reduce_last_axis": [C.c_void_p, C.c_int32, C.c_int32, _P, C.c_int64, _P],
    "shard_arg_reduce_last_axis": [C.c_void_p, C.c_int32, C.c_int32, _P, C.c_int64, _P],
    "shard_find_last_axis": [C.c_void_p, C.c_int32, C.c_void_p, _P, C.c_int64, _P],
    "shard_minmax_arg_last_axis": [C.c_void_p, C.c_int32, C.c_int32, _P, _P, C.c_int64, _P],
    "shard_reduce_sharded_axis": [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, _P, C.c_int64, _P],
    "shard_all_gather_rows": [C.c_void_p, C.c_int32, _P, C.c_int64, C.c_int64],
    "shard_count_true": [C.c_void_p, C.c_int32, _P, C.POINTER(C.c_int64)],
    "shard_count_true_begin": [C.c_void_p, C.c_int32, _P],
    "shard_count_true_end": [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)],
    "shard_true_indices": [C.c_void_p, C.c_int32, _P, C.c_int64, C.c_int64, _P, C.c_int64],
    "shard_masked_get": [C.c_void_p, C.c_int32, _P, C.c_int64, C.c_int64, _P, _P],
}
DN_SHARD_MAX_RANKS, DN_SHARD_HANDLE_BYTES = 8, 64

ALL_PRODUCT_SYMBOLS = (
    ["dn_" + n for n in _OPERATOR_SIGNATURES] + ["dn_" + n for n in _DEVICE_SIGNATURES] +
    ["dn_last_error", "dn_launch_count", "dn_version"]
)


class CApi:
    """One loaded shared library exposing the operator entry points under `prefix`."""

    def __init__(self, path: str, prefix: str, with_device_api: bool):
        if not os.path.exists(path):
            raise ImportError(
                f"native library {path} is missing — build it first (python -c 'import __graft_entry__ as g; "
                f"g.build()' or `make`). There is no CPU fallback.")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        sigs = dict(_OPERATOR_SIGNATURES)
        if with_device_api:
            sigs.update(_DEVICE_SIGNATURES)
        for name, argtypes in sigs.items():
            fn = getattr(self.lib, prefix + name)
            fn.argtypes = argtypes
            fn.restype = C.c_int32
            setattr(self, "_" + name, fn)
        self._last_error = getattr(self.lib, prefix + "last_error")
        self._last_error.restype = C.c_char_p
        self._last_error.argtypes = []
        if with_device_api:
            self.lib.dn_launch_count.restype = C.c_int64
            self.lib.dn_launch_count.argtypes = []
            self.lib.dn_version.restype = C.c_char_p
            self.lib.dn_version.argtypes = []

    def check(self, status: int) -> None:
        if status != DN_OK:
            msg = (self._last_error() or b"").decode("utf-8", "replace")
            raise _EXC.get(status, RuntimeError)(msg or f"native call failed with status {status}")

    def call(self, name: str, *args) -> None:
        self.check(getattr(self, "_" + name)(*args))


def scalar_buffer(value, dtype: int):
    """One host element of `dtype` as a ctypes-passable buffer (scalars are passed by pointer)."""
    arr = np.array([value], dtype=dtypes.to_numpy(dtype))
    return arr, arr.ctypes.data_as(C.c_void_p)


def desc_ptr_array(descs: Sequence[Optional[dn_tensor]]):
    """`const dn_tensor* const*` with NULL for None entries."""
    arr = (_P * max(1, len(descs)))()
    for i, d in enumerate(descs):
        arr[i] = C.pointer(d) if d is not None else None
    return arr


_PRODUCT: Optional[CApi] = None


def product_library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdeepnet_b200.so")


def product() -> CApi:
    """The product library. Fails loudly if it has not been built — there is no fallback path."""
    global _PRODUCT
    if _PRODUCT is None:
        _PRODUCT = CApi(product_library_path(), "dn_", with_device_api=True)
    return _PRODUCT
