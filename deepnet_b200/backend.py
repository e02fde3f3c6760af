"""ITensorDevice / ITensorStorage / ITensorBackend and their CUDA (B200) implementation.

Mirrors the reference's backend boundary, Tensor/Tensor/TensorBackend.fs:14-146, member for member: every
`ITensorBackend` method below has the reference's name and argument order (target first, sources pre-broadcast to
the target's shape by the frontend). `TensorCudaDevice` / `TensorCudaStorage` / `TensorCudaBackend` replace
Tensor/Tensor/Cuda/CudaBackend.fs:51-108,120-492,497-507; all they do is marshal layouts into `dn_tensor`
descriptors and call libdeepnet_b200.so — there is no kernel table, no NVRTC and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import dtypes, native
from . import layout as TL
from .layout import TensorLayout
from .native import CApi, NotSupportedException

# dn_unary_op / dn_binary_op / dn_compare_op / dn_reduce_op / dn_arg_reduce_op (include/dn_tensor.h)
UNARY_OPS = ["UnaryPlus", "UnaryMinus", "Abs", "Sgn", "Log", "Log10", "Exp", "Sin", "Cos", "Tan", "Asin", "Acos",
             "Atan", "Sinh", "Cosh", "Tanh", "Sqrt", "Ceiling", "Floor", "Round", "Truncate", "Negate"]
BINARY_OPS = ["Add", "Subtract", "Multiply", "Divide", "Modulo", "Power", "MaxElemwise", "MinElemwise",
              "And", "Or", "Xor"]
COMPARE_OPS = ["Equal", "NotEqual", "Less", "LessOrEqual", "Greater", "GreaterOrEqual"]
REDUCE_OPS = ["SumLastAxis", "ProductLastAxis", "MinLastAxis", "MaxLastAxis", "AllLastAxis", "AnyLastAxis",
              "CountTrueLastAxis"]
ARG_REDUCE_OPS = ["ArgMinLastAxis", "ArgMaxLastAxis"]


class ITensorDevice:
    """TensorBackend.fs:23-30."""
    Id: str = ""
    Zeroed: bool = False

    def Create(self, nElems: int, dtype: int) -> "ITensorStorage":
        raise NotImplementedError

    def __eq__(self, other):
        return isinstance(other, ITensorDevice) and self.Id == other.Id

    def __hash__(self):
        return hash(self.Id)

    def __repr__(self):
        return self.Id


class ITensorStorage:
    """TensorBackend.fs:14-20. `DataType` is the dn_dtype of 'T."""
    Dev: ITensorDevice
    DataType: int

    def Backend(self, layout: TensorLayout) -> "NativeTensorBackend":
        raise NotImplementedError

    def BasePtr(self) -> int:
        raise NotImplementedError


class NativeTensorBackend:
    """ITensorBackend<'T> (TensorBackend.fs:64-146) over a C ABI library.

    Arguments are frontends (anything with `.Storage`, `.Layout`, `.DataType`); this class only builds descriptors
    (the role of TensorCudaBackend.GetNativeTensor, CudaBackend.fs:151-198) and calls the entry point.
    """

    def __init__(self, layout: TensorLayout, storage: ITensorStorage, api: CApi):
        self.layout = layout
        self.storage = storage
        self.api = api

    # -- descriptor marshalling ------------------------------------------------------------------------------
    @staticmethod
    def _d(t) -> native.dn_tensor:
        """The dn_tensor descriptor of a frontend. A Tensor's layout and storage never change after construction,
        and the native side treats descriptors as const, so the descriptor is built once per Tensor object."""
        d = getattr(t, "_desc", None)
        if d is None:
            d = native.make_desc(t.Storage.BasePtr(), t.Layout, t.DataType)
            try:
                t._desc = d
            except AttributeError:
                pass
        return d

    def _call(self, name, *args):
        self.api.call(name, *args)

    # -- Item (TensorBackend.fs:65) ------------------------------------------------------------------------
    def GetItem(self, idx: Sequence[int]):
        raise NotImplementedError

    def SetItem(self, idx: Sequence[int], value) -> None:
        raise NotImplementedError

    # -- Copy / Convert / Fill (TensorBackend.fs:67-72) ----------------------------------------------------
    def Copy(self, trgt, src) -> None:
        self._call("copy", self._d(trgt), self._d(src))

    def Transfer(self, trgt, src) -> bool:
        return False

    def Convert(self, trgt, src) -> None:
        self._call("convert", self._d(trgt), self._d(src))

    def FillConst(self, value, trgt) -> None:
        keep, p = native.scalar_buffer(value, trgt.DataType)
        self._call("fill_const", self._d(trgt), p)

    def FillIncrementing(self, start, incr, trgt) -> None:
        k1, p1 = native.scalar_buffer(start, trgt.DataType)
        k2, p2 = native.scalar_buffer(incr, trgt.DataType)
        self._call("fill_incrementing", self._d(trgt), p1, p2)

    # -- element-wise (TensorBackend.fs:74-118) ------------------------------------------------------------
    def _unary(self, op: int, trgt, src1) -> None:
        self._call("unary", op, self._d(trgt), self._d(src1))

    def _binary(self, op: int, trgt, src1, src2) -> None:
        self._call("binary", op, self._d(trgt), self._d(src1), self._d(src2))

    def _compare(self, op: int, trgt, src1, src2) -> None:
        self._call("compare", op, self._d(trgt), self._d(src1), self._d(src2))

    def IsFinite(self, trgt, src1) -> None:
        self._call("is_finite", self._d(trgt), self._d(src1))

    def IfThenElse(self, trgt, cond, ifTrue, ifFalse) -> None:
        self._call("if_then_else", self._d(trgt), self._d(cond), self._d(ifTrue), self._d(ifFalse))

    # -- indexing (TensorBackend.fs:119-123) ---------------------------------------------------------------
    def Gather(self, trgt, srcIdxs: List[Optional[object]], src) -> None:
        descs = [self._d(i) if i is not None else None for i in srcIdxs]
        self._call("gather", self._d(trgt), native.desc_ptr_array(descs), len(descs), self._d(src))

    def Scatter(self, trgt, trgtIdxs: List[Optional[object]], src) -> None:
        descs = [self._d(i) if i is not None else None for i in trgtIdxs]
        self._call("scatter", self._d(trgt), native.desc_ptr_array(descs), len(descs), self._d(src))

    def MaskedGet(self, trgt, src, masks: List[Optional[object]]) -> None:
        descs = [self._d(m) if m is not None else None for m in masks]
        self._call("masked_get", self._d(trgt), self._d(src), native.desc_ptr_array(descs), len(descs))

    def MaskedSet(self, trgt, masks: List[Optional[object]], src) -> None:
        descs = [self._d(m) if m is not None else None for m in masks]
        self._call("masked_set", self._d(trgt), native.desc_ptr_array(descs), len(descs), self._d(src))

    def TrueIndices(self, trgt, src1) -> None:
        self._call("true_indices", self._d(trgt), self._d(src1))

    def CountTrue(self, src1) -> int:
        """countTrue ∘ flatten with the blocking read-back the frontend needs (Tensor.fs:2232-2233,2259-2262)."""
        n = C.c_int64(0)
        self._call("count_true", self._d(src1), C.byref(n))
        return int(n.value)

    # -- reductions (TensorBackend.fs:125-135) -------------------------------------------------------------
    def _reduce(self, op: int, trgt, src1) -> None:
        self._call("reduce_last_axis", op, self._d(trgt), self._d(src1))

    def _arg_reduce(self, op: int, trgt, src1) -> None:
        self._call("arg_reduce_last_axis", op, self._d(trgt), self._d(src1))

    def FindLastAxis(self, value, trgt, src1) -> None:
        keep, p = native.scalar_buffer(value, src1.DataType)
        self._call("find_last_axis", p, self._d(trgt), self._d(src1))

    # -- dot products (TensorBackend.fs:137-140) -----------------------------------------------------------
    def VecVecDot(self, trgt, src1, src2) -> None:
        self._call("vec_vec_dot", self._d(trgt), self._d(src1), self._d(src2))

    def MatVecDot(self, trgt, src1, src2) -> None:
        self._call("mat_vec_dot", self._d(trgt), self._d(src1), self._d(src2))

    def MatMatDot(self, trgt, src1, src2) -> None:
        self._call("mat_mat_dot", self._d(trgt), self._d(src1), self._d(src2))

    def BatchedMatMatDot(self, trgt, src1, src2) -> None:
        self._call("batched_mat_mat_dot", self._d(trgt), self._d(src1), self._d(src2))

    def FusedElemwise(self, trgt, srcs: List[object], prog: List[tuple]) -> None:
        """dn_fused_elemwise (new): `prog` = [(kind, op, dst, a, b, imm)], see include/dn_tensor.h."""
        arr = (native.dn_fused_instr * len(prog))()
        for i, (kind, op, dst, a, b, imm) in enumerate(prog):
            arr[i] = native.dn_fused_instr(kind, op, dst, a, b, float(imm))
        descs = [self._d(x) for x in srcs]
        self._call("fused_elemwise", self._d(trgt), native.desc_ptr_array(descs), len(descs), arr, len(prog))

    def BatchedInvert(self, trgt, src) -> None:
        """TensorBackend.fs:142; CudaBackend.fs:451-484. Raises SingularMatrixException."""
        self._call("batched_invert", self._d(trgt), self._d(src))

    # -- not part of the hot path; unsupported on CUDA in the reference too (CudaBackend.fs:486-488) --------
    def BatchedSVD(self, *a):
        raise NotSupportedException("the CUDA tensor backend currently does not support the BatchedSVD operation")

    def SymmetricEigenDecomposition(self, *a):
        raise NotSupportedException(
            "the CUDA tensor backend currently does not support the SymmetricEigenDecomposition operation")


def _install_named_members():
    """Generate UnaryPlus..Negate, Add..Xor, Equal..GreaterOrEqual, *LastAxis with the reference's names."""
    def mk_unary(code):
        return lambda self, trgt, src1: self._unary(code, trgt, src1)

    def mk_binary(code):
        return lambda self, trgt, src1, src2: self._binary(code, trgt, src1, src2)

    def mk_compare(code):
        return lambda self, trgt, src1, src2: self._compare(code, trgt, src1, src2)

    def mk_reduce(code):
        return lambda self, trgt, src1: self._reduce(code, trgt, src1)

    def mk_arg(code):
        return lambda self, trgt, src1: self._arg_reduce(code, trgt, src1)

    for table, mk in ((UNARY_OPS, mk_unary), (BINARY_OPS, mk_binary), (COMPARE_OPS, mk_compare),
                      (REDUCE_OPS, mk_reduce), (ARG_REDUCE_OPS, mk_arg)):
        for code, name in enumerate(table):
            fn = mk(code)
            fn.__name__ = name
            setattr(NativeTensorBackend, name, fn)


_install_named_members()


# =================================================================================================================
# CUDA device (B200)
# =================================================================================================================
class TensorCudaBackend(NativeTensorBackend):
    """Replaces TensorCudaBackend<'T>, CudaBackend.fs:120-492."""

    def GetItem(self, idx):
        out = np.zeros(1, dtype=dtypes.to_numpy(self.storage.DataType))
        pos = (C.c_int64 * max(1, len(idx)))(*idx)
        d = native.make_desc(self.storage.BasePtr(), self.layout, self.storage.DataType)
        self._call("get_item", d, pos, out.ctypes.data_as(C.c_void_p))
        return out[0]

    def SetItem(self, idx, value):
        keep, p = native.scalar_buffer(value, self.storage.DataType)
        pos = (C.c_int64 * max(1, len(idx)))(*idx)
        d = native.make_desc(self.storage.BasePtr(), self.layout, self.storage.DataType)
        self._call("set_item", d, pos, p)

    def Transfer(self, trgt, src) -> bool:
        """CudaBackend.fs:206-270. Any pair of views: the layout handling (strided pack of the host side, strided
        copy on the device) is native — dn_transfer_h2d / dn_transfer_d2h. Returns False when this backend cannot
        do the transfer (both or neither side on the device)."""
        t_cuda = isinstance(trgt.Storage, TensorCudaStorage)
        s_cuda = isinstance(src.Storage, TensorCudaStorage)
        if t_cuda == s_cuda:
            return False
        if t_cuda:
            self._call("transfer_h2d", self._d(trgt), self._d(src))
        else:
            self._call("transfer_d2h", self._d(trgt), self._d(src))
        return True


class TensorCudaStorage(ITensorStorage):
    """Replaces TensorCudaStorage<'T>, CudaBackend.fs:51-108. Owns stream-ordered device memory, or wraps an
    external device pointer without owning it (CudaFrontend.fs:129-137 `usingPtr`)."""

    def __init__(self, nElems: int, dtype: int, dev: "TensorCudaDevice", ptr: Optional[int] = None, owner=None):
        self.Dev = dev
        self.DataType = dtype
        self.DataSize = max(1, int(nElems))  # CUDA cannot allocate size zero, CudaBackend.fs:56-58
        self.DataSizeInBytes = self.DataSize * dtypes.itemsize(dtype)
        self._api = dev.api
        self._owner = owner
        if ptr is None:
            p = C.c_void_p()
            self._api.call("alloc", self.DataSizeInBytes, C.byref(p))
            self._ptr = p.value
            self._owned = True
        else:
            self._ptr = int(ptr)
            self._owned = False

    def BasePtr(self) -> int:
        return self._ptr

    def Backend(self, layout: TensorLayout) -> TensorCudaBackend:
        return TensorCudaBackend(layout, self, self._api)

    def __del__(self):  # finalizer frees, CudaBackend.fs:73-74; the free is stream-ordered (dn_free)
        try:
            if getattr(self, "_owned", False) and self._ptr:
                self._api._free(self._ptr)
                self._ptr = 0
        except Exception:
            pass


class TensorCudaDevice(ITensorDevice):
    """Replaces TensorCudaDevice, CudaBackend.fs:497-507 (`Id = "Cuda"`, `Zeroed = false`)."""
    Id = "Cuda"
    Zeroed = False
    _instance: Optional["TensorCudaDevice"] = None

    def __init__(self):
        self.api = native.product()

    @classmethod
    def Instance(cls) -> "TensorCudaDevice":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def Init(self, device: int = 0) -> None:
        """CudaInit.check, CudaBackend.fs:28-38."""
        self.api.call("init", device)

    def Create(self, nElems: int, dtype: int) -> TensorCudaStorage:
        return TensorCudaStorage(nElems, dtype, self)

    def UsingPtr(self, ptr: int, nElems: int, dtype: int, owner=None) -> TensorCudaStorage:
        """CudaTensor.usingPtr, CudaFrontend.fs:129-137."""
        return TensorCudaStorage(nElems, dtype, self, ptr=ptr, owner=owner)

    def Synchronize(self) -> None:
        self.api.call("sync")

    def SetStream(self, stream: int) -> None:
        """Cfg.Stream, CudaCfg.fs:25-27."""
        self.api.call("set_stream", stream)

    def GetStream(self) -> int:
        import ctypes
        st = ctypes.c_void_p()
        self.api.call("get_stream", ctypes.byref(st))
        return int(st.value or 0)

    def SetStacktrace(self, enabled: bool) -> None:
        """Cfg.Stacktrace, CudaCfg.fs:33-35."""
        self.api.call("set_check_errors", 1 if enabled else 0)

    def SetMathMode(self, mode: str) -> None:
        """dn_set_math_mode: "fp32" (default: exact kernel for small products, 3xTF32 on the tensor cores for large
        ones), "tf32" (one tf32 pass) or "strict" (exact fp32 kernel for every size)."""
        self.api.call("set_math_mode", {"fp32": 0, "tf32": 1, "strict": 2}[mode])

    def LaunchCount(self) -> int:
        return int(self.api.lib.dn_launch_count())


# =================================================================================================================
# Pinned-host staging device: plain host memory holding tensors on their way to / from the GPU. It has storage but
# NO compute backend — operators on it raise, by design ("no CPU fallback", BASELINE.json north_star).
# =================================================================================================================
class TensorStagingStorage(ITensorStorage):
    def __init__(self, array: np.ndarray, dev: "TensorStagingDevice"):
        self.Dev = dev
        self.array = array  # 1-D, owns or aliases host memory
        self.DataType = dtypes.from_numpy(array.dtype)
        self.DataSize = array.size

    def BasePtr(self) -> int:
        return self.array.ctypes.data

    def Backend(self, layout):
        raise NotSupportedException(
            "host staging tensors have no compute backend: transfer to CudaTensor.Dev first (there is no CPU "
            "fallback in deepnet_b200)")


class TensorStagingDevice(ITensorDevice):
    Id = "HostStaging"
    Zeroed = True
    _instance = None

    @classmethod
    def Instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def Create(self, nElems, dtype):
        return TensorStagingStorage(np.zeros(max(1, int(nElems)), dtype=dtypes.to_numpy(dtype)), self)
