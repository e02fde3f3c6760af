"""Element types of the C ABI (include/dn_tensor.h `dn_dtype`) and their numpy counterparts.

The reference maps .NET primitive types to C++ type names in Tensor/Tensor/Cuda/NativeTensor.fs:22-34; the order
below is ABI and must match `dn_dtype`.
"""
import numpy as np

DN_F32, DN_F64, DN_I8, DN_U8, DN_I16, DN_U16, DN_I32, DN_U32, DN_I64, DN_U64, DN_BOOL = range(11)

_NP = {
    DN_F32: np.float32, DN_F64: np.float64, DN_I8: np.int8, DN_U8: np.uint8, DN_I16: np.int16,
    DN_U16: np.uint16, DN_I32: np.int32, DN_U32: np.uint32, DN_I64: np.int64, DN_U64: np.uint64,
    DN_BOOL: np.bool_,
}
_FROM_NP = {np.dtype(v): k for k, v in _NP.items()}

NAMES = {
    DN_F32: "single", DN_F64: "double", DN_I8: "sbyte", DN_U8: "byte", DN_I16: "int16", DN_U16: "uint16",
    DN_I32: "int32", DN_U32: "uint32", DN_I64: "int64", DN_U64: "uint64", DN_BOOL: "bool",
}


def to_numpy(dtype: int) -> np.dtype:
    return np.dtype(_NP[dtype])


def from_numpy(np_dtype) -> int:
    try:
        return _FROM_NP[np.dtype(np_dtype)]
    except KeyError:
        raise TypeError(f"unsupported element type {np_dtype}") from None


def itemsize(dtype: int) -> int:
    return to_numpy(dtype).itemsize


def is_float(dtype: int) -> bool:
    return dtype in (DN_F32, DN_F64)
