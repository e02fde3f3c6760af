"""Tensor — the device-agnostic frontend, mirroring the parts of Tensor/Tensor/Tensor.fs that feed the backend.

The reference's `Tensor<'T>` does shape checks, broadcasting (stride-0 views), target allocation and the
axis→last-axis permutation, then calls `trgt.Backend.Op(trgt, srcs…)`. This module restates
exactly that host logic (file:line cited per member) so that code written against the reference's API shape runs
unchanged on `CudaTensor.Dev`; all arithmetic happens behind `ITensorBackend` in libdeepnet_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import dtypes
from . import layout as TL
from .backend import ITensorDevice, ITensorStorage, TensorCudaDevice, TensorStagingDevice, TensorStagingStorage
from .layout import NotFound, TensorLayout

NoMask = None  # SpecialMask.NoMask, Tensor.fs:4604-4607


class Tensor:
    """Tensor<'T>, Tensor.fs:50-54: a layout over a storage; the backend is created once per tensor object."""

    def __init__(self, layout: TensorLayout, storage: ITensorStorage):
        self._layout = layout
        self._storage = storage
        self._backend = None

    # ---- ITensorFrontend (TensorBackend.fs:33-61) ----------------------------------------------------------
    @property
    def Layout(self) -> TensorLayout: return self._layout
    @property
    def Storage(self) -> ITensorStorage: return self._storage
    @property
    def Dev(self) -> ITensorDevice: return self._storage.Dev
    @property
    def DataType(self) -> int: return self._storage.DataType
    @property
    def Shape(self) -> Tuple[int, ...]: return self._layout.Shape
    @property
    def Stride(self) -> Tuple[int, ...]: return self._layout.Stride
    @property
    def Offset(self) -> int: return self._layout.Offset
    @property
    def NDims(self) -> int: return self._layout.NDims
    @property
    def NElems(self) -> int: return self._layout.NElems

    @property
    def Backend(self):
        if self._backend is None:
            self._backend = self._storage.Backend(self._layout)  # Tensor.fs:54
        return self._backend

    def Relayout(self, layout: TensorLayout) -> "Tensor":
        return Tensor(layout, self._storage)

    # ---- construction (Tensor.fs:317-325, 3775-3872) --------------------------------------------------------
    @staticmethod
    def empty(shape: Sequence[int], dtype: int, dev: ITensorDevice, order: str = "C") -> "Tensor":
        shape = tuple(int(s) for s in shape)
        lay = TL.newC(shape) if order == "C" else TL.newF(shape)
        return Tensor(lay, dev.Create(lay.NElems, dtype))

    @staticmethod
    def filled(shape, value, dtype, dev) -> "Tensor":
        t = Tensor.empty(shape, dtype, dev)
        t.FillConst(value)
        return t

    @staticmethod
    def zeros(shape, dtype, dev) -> "Tensor":
        return Tensor.filled(shape, 0, dtype, dev)

    @staticmethod
    def ones(shape, dtype, dev) -> "Tensor":
        return Tensor.filled(shape, 1, dtype, dev)

    @staticmethod
    def scalar(value, dtype, dev) -> "Tensor":
        return Tensor.filled((), value, dtype, dev)

    @staticmethod
    def counting(dev, nElems: int) -> "Tensor":
        """Tensor.counting, Tensor.fs:3816-3831."""
        t = Tensor.empty((nElems,), dtypes.DN_I64, dev)
        t.FillIncrementing(0, 1)
        return t

    @staticmethod
    def arange(dev, start, incr, stop, dtype=dtypes.DN_F64) -> "Tensor":
        """Tensor.arange, Tensor.fs:3834-3857: nElems = max 0 (int64 ((stop - start) / incr))."""
        n = max(0, int((stop - start) / incr))
        t = Tensor.empty((n,), dtype, dev)
        t.FillIncrementing(start, incr)
        return t

    @staticmethod
    def linspace(dev, start, stop, nElems: int, dtype=dtypes.DN_F64) -> "Tensor":
        """Tensor.linspace, Tensor.fs:3860-3886: increment (stop - start) / (nElems - 1)."""
        if nElems < 2:
            raise ValueError("linspace requires at least two elements")
        incr = (stop - start) / (nElems - 1)
        t = Tensor.empty((nElems,), dtype, dev)
        t.FillIncrementing(start, incr)
        return t

    @staticmethod
    def ofNumpy(arr: np.ndarray, dev: Optional[ITensorDevice] = None) -> "Tensor":
        """Wrap (a C-contiguous copy of) a numpy array as a host staging tensor, or transfer it to `dev`."""
        arr = np.ascontiguousarray(arr)
        st = TensorStagingStorage(arr.reshape(-1) if arr.size else np.zeros(1, arr.dtype),
                                  TensorStagingDevice.Instance())
        t = Tensor(TL.newC(arr.shape), st)
        return t if dev is None or dev == t.Dev else Tensor.transfer(dev, t)

    def toNumpy(self) -> np.ndarray:
        """Materialise as a numpy array with this tensor's logical contents (blocking for device tensors)."""
        st = self._storage
        if not hasattr(st, "array"):
            host = Tensor.transfer(TensorStagingDevice.Instance(), self)
            return host.toNumpy()
        isz = st.array.itemsize
        if self.NElems == 0:
            return np.zeros(self.Shape, dtype=st.array.dtype)
        v = np.lib.stride_tricks.as_strided(st.array[self.Offset:], shape=self.Shape,
                                            strides=tuple(s * isz for s in self.Stride), writeable=False)
        return np.array(v)

    # ---- transfer (Tensor.fs:3706-3745; CudaBackend.fs:206-270) --------------------------------------------
    def TransferFrom(self, src: "Tensor") -> None:
        if self.Shape != src.Shape or self.DataType != src.DataType:
            raise ValueError(f"cannot transfer tensor of shape {src.Shape} into tensor of shape {self.Shape}")
        for side in (self, src):
            try:
                be = side.Backend
            except NotImplementedError:
                continue
            if be.Transfer(self, src):
                return
        raise RuntimeError(f"Cannot transfer from storage {src.Dev} to storage {self.Dev}.")

    @staticmethod
    def transfer(dev: ITensorDevice, src: "Tensor") -> "Tensor":
        if src.Dev == dev:
            return src
        trgt = Tensor.empty(src.Shape, src.DataType, dev)
        trgt.TransferFrom(src)
        return trgt

    # ---- views (Tensor.fs:330-640; TensorLayout.fs) --------------------------------------------------------
    @property
    def T(self) -> "Tensor": return self.Relayout(TL.transpose(self._layout))

    def __getitem__(self, rngs) -> "Tensor":
        if not isinstance(rngs, tuple):
            rngs = (rngs,)
        return self.Relayout(TL.view(rngs, self._layout))

    def __setitem__(self, rngs, value: "Tensor") -> None:
        """SetRng, Tensor.fs:2990-2993."""
        trgt = self[rngs]
        trgt.CopyFrom(value.broadcastTo(trgt.Shape))

    def swapDim(self, ax1, ax2): return self.Relayout(TL.swapDim(ax1, ax2, self._layout))
    def permuteAxes(self, permut): return self.Relayout(TL.permuteAxes(permut, self._layout))
    def reverseAxis(self, ax): return self.Relayout(TL.reverseAxis(ax, self._layout))
    def broadcastTo(self, shp): return self.Relayout(TL.broadcastToShape(shp, self._layout))
    def padLeft(self): return self.Relayout(TL.padLeft(self._layout))
    def padRight(self): return self.Relayout(TL.padRight(self._layout))
    def diagAxis(self, ax1, ax2): return self.Relayout(TL.diagAxis(ax1, ax2, self._layout))

    def diag(self) -> "Tensor":
        if self.NDims < 2:
            raise ValueError("Need at least a matrix to extract diagonal")
        return self.diagAxis(self.NDims - 2, self.NDims - 1)

    @staticmethod
    def diagMatAxis(ax1: int, ax2: int, a: "Tensor") -> "Tensor":
        """Tensor.diagMatAxis, Tensor.fs:4140-4151: a new tensor with `a` on the diagonal over (ax1, ax2), zero elsewhere."""
        if ax1 == ax2:
            raise ValueError("axes to use for diagonal must be different")
        ax1, ax2 = (ax1, ax2) if ax1 < ax2 else (ax2, ax1)
        if not (0 <= ax1 < a.NDims):
            raise ValueError(f"Specified axis {ax1} is invalid for tensor of shape {a.Shape}.")
        if not (0 <= ax2 <= a.NDims):
            raise ValueError(f"Cannot insert axis at position {ax2} into array of shape {a.Shape}.")
        shp = list(a.Shape)
        shp.insert(ax2, a.Shape[ax1])
        d = Tensor.zeros(shp, a.DataType, a.Dev)
        d.diagAxis(ax1, ax2).CopyFrom(a)
        return d

    @staticmethod
    def diagMat(a: "Tensor") -> "Tensor":
        """Tensor.diagMat, Tensor.fs:4168-4171."""
        if a.NDims < 1:
            raise ValueError("need at leat a one-dimensional array to create a diagonal matrix")
        return Tensor.diagMatAxis(a.NDims - 1, a.NDims, a)

    def traceAxis(self, ax1: int, ax2: int) -> "Tensor":
        """Tensor.traceAxis, Tensor.fs:4188-4190: sum over the diagonal view."""
        tax = ax1 if ax1 < ax2 else ax1 - 1
        return self.diagAxis(ax1, ax2).sumAxis(tax)

    def trace(self) -> "Tensor":
        """Tensor.trace, Tensor.fs:4206-4209."""
        if self.NDims < 2:
            raise ValueError(f"Need at least a two dimensional array for trace but got shape {self.Shape}.")
        return self.traceAxis(self.NDims - 2, self.NDims - 1)

    def tryReshapeView(self, shp) -> Optional["Tensor"]:
        lay = TL.tryReshape(shp, self._layout)
        return None if lay is None else self.Relayout(lay)

    def reshape(self, shp) -> "Tensor":
        """Tensor.reshape, Tensor.fs:466-480: view if possible, otherwise copy first."""
        v = self.tryReshapeView(shp)
        return v if v is not None else self.Copy().tryReshapeView(shp)

    def flatten(self) -> "Tensor":
        return self.reshape((TL.Remainder,))

    @staticmethod
    def broadcastToSame(*xs: "Tensor") -> List["Tensor"]:
        lays = TL.broadcastToSameMany([x.Layout for x in xs])
        return [x.Relayout(l) for x, l in zip(xs, lays)]

    # ---- copy / convert / fill (Tensor.fs:640-830) ------------------------------------------------------------
    def CopyFrom(self, src: "Tensor") -> None:
        Tensor.CheckSameStorage(self, src)
        if self.Shape != src.Shape:
            raise ValueError(f"Tensors of shapes {self.Shape} and {src.Shape} were expected to have same shape")
        self.Backend.Copy(self, src)

    def Copy(self, order: str = "C") -> "Tensor":
        trgt = Tensor.empty(self.Shape, self.DataType, self.Dev, order=order)
        trgt.CopyFrom(self)
        return trgt

    def FillConst(self, value) -> None: self.Backend.FillConst(value, self)
    def FillIncrementing(self, start, incr) -> None: self.Backend.FillIncrementing(start, incr, self)

    def FillConvert(self, a: "Tensor") -> None:
        a = Tensor.PrepareElemwiseSources(self, a)[0]
        self.Backend.Convert(self, a)

    def convert(self, dtype: int) -> "Tensor":
        trgt = Tensor.empty(self.Shape, dtype, self.Dev)
        trgt.FillConvert(self)
        return trgt

    def Item(self, *idx):
        return self.Backend.GetItem(list(idx))

    def SetItem(self, idx, value):
        self.Backend.SetItem(list(idx), value)

    @property
    def Value(self):
        """Tensor.value, Tensor.fs:3395-3398 (blocking on device tensors)."""
        if self.NDims != 0:
            raise ValueError(f"Value is only available for scalar tensors, but shape is {self.Shape}")
        return self.Backend.GetItem([])

    # ---- helpers (Tensor.fs:4507-4597) ----------------------------------------------------------------------
    @staticmethod
    def CheckSameStorage(*xs: "Tensor") -> None:
        if any(x.Dev != xs[0].Dev for x in xs[1:]):
            raise RuntimeError(
                f"Storage devices must be equal for this operation, but they are {[x.Dev.Id for x in xs]}.")

    @staticmethod
    def PrepareElemwiseSources(trgt: "Tensor", *srcs: "Tensor") -> List["Tensor"]:
        Tensor.CheckSameStorage(trgt, *srcs)
        return [s.broadcastTo(trgt.Shape) for s in srcs]

    @staticmethod
    def PrepareElemwise(dtype: int, *srcs: "Tensor"):
        Tensor.CheckSameStorage(*srcs)
        bc = Tensor.broadcastToSame(*srcs)
        trgt = Tensor.empty(bc[0].Shape, dtype, bc[0].Dev)
        return (trgt, *bc)

    @staticmethod
    def PrepareAxisReduceSources(trgt: "Tensor", axis: int, a: "Tensor") -> "Tensor":
        Tensor.CheckSameStorage(trgt, a)
        TL.checkAxis(axis, a.Layout)
        red = a.Shape[:axis] + a.Shape[axis + 1:]
        if trgt.Shape != red:
            raise RuntimeError(f"Reduction of tensor {a.Shape} along axis {axis} gives shape {red} but target "
                               f"has shape {trgt.Shape}.")
        axis_to_last = list(range(axis)) + [a.NDims - 1] + list(range(axis, a.NDims - 1))
        return a.permuteAxes(axis_to_last)

    @staticmethod
    def PrepareAxisReduceTarget(dtype: int, axis: int, a: "Tensor") -> "Tensor":
        TL.checkAxis(axis, a.Layout)
        return Tensor.empty(a.Shape[:axis] + a.Shape[axis + 1:], dtype, a.Dev)

    def _coerce(self, other) -> "Tensor":
        """Scalar operands become rank-0 tensors on the same device (scalarLike, Tensor.fs:1372)."""
        if isinstance(other, Tensor):
            return other
        return Tensor.scalar(other, self.DataType, self.Dev)

    # ---- element-wise operators: Fill* + allocating forms (Tensor.fs:836-2083) -------------------------------
    def _fill_unary(self, member: str, a: "Tensor") -> None:
        a, = Tensor.PrepareElemwiseSources(self, a)
        getattr(self.Backend, member)(self, a)

    def _fill_binary(self, member: str, a: "Tensor", b: "Tensor") -> None:
        a, b = Tensor.PrepareElemwiseSources(self, a, b)
        getattr(self.Backend, member)(self, a, b)

    def _new_unary(self, member: str, dtype: Optional[int] = None) -> "Tensor":
        trgt, a = Tensor.PrepareElemwise(self.DataType if dtype is None else dtype, self)
        # comparison-like ops dispatch on the SOURCE's backend (Tensor.fs:1736)
        getattr(a.Backend if dtype is not None else trgt.Backend, member)(trgt, a)
        return trgt

    def _new_binary(self, member: str, other, dtype: Optional[int] = None) -> "Tensor":
        other = self._coerce(other)
        trgt, a, b = Tensor.PrepareElemwise(self.DataType if dtype is None else dtype, self, other)
        getattr(a.Backend if dtype is not None else trgt.Backend, member)(trgt, a, b)
        return trgt

    def __pos__(self): return self._new_unary("UnaryPlus")
    def __neg__(self): return self._new_unary("UnaryMinus")
    def __abs__(self): return self._new_unary("Abs")
    def __add__(self, o): return self._new_binary("Add", o)
    def __radd__(self, o): return self._coerce(o)._new_binary("Add", self)
    def __sub__(self, o): return self._new_binary("Subtract", o)
    def __rsub__(self, o): return self._coerce(o)._new_binary("Subtract", self)
    def __mul__(self, o): return self._new_binary("Multiply", o)
    def __rmul__(self, o): return self._coerce(o)._new_binary("Multiply", self)
    def __truediv__(self, o): return self._new_binary("Divide", o)
    def __rtruediv__(self, o): return self._coerce(o)._new_binary("Divide", self)
    def __mod__(self, o): return self._new_binary("Modulo", o)
    def __pow__(self, o): return self._new_binary("Power", o)
    def __invert__(self): return self._new_unary("Negate")
    def __and__(self, o): return self._new_binary("And", o)
    def __or__(self, o): return self._new_binary("Or", o)
    def __xor__(self, o): return self._new_binary("Xor", o)
    # element-wise comparisons: ==== <<<< etc. in F# (Tensor.fs:1722-1985); Python spells them eq/ne/lt/...
    def eq(self, o): return self._new_binary("Equal", o, dtypes.DN_BOOL)
    def ne(self, o): return self._new_binary("NotEqual", o, dtypes.DN_BOOL)
    def lt(self, o): return self._new_binary("Less", o, dtypes.DN_BOOL)
    def le(self, o): return self._new_binary("LessOrEqual", o, dtypes.DN_BOOL)
    def gt(self, o): return self._new_binary("Greater", o, dtypes.DN_BOOL)
    def ge(self, o): return self._new_binary("GreaterOrEqual", o, dtypes.DN_BOOL)
    def isFinite(self): return self._new_unary("IsFinite", dtypes.DN_BOOL)

    @staticmethod
    def maxElemwise(a: "Tensor", b) -> "Tensor": return a._new_binary("MaxElemwise", b)
    @staticmethod
    def minElemwise(a: "Tensor", b) -> "Tensor": return a._new_binary("MinElemwise", b)

    def FillIfThenElse(self, cond: "Tensor", ifTrue: "Tensor", ifFalse: "Tensor") -> None:
        cond, ifTrue, ifFalse = Tensor.PrepareElemwiseSources(self, cond, ifTrue, ifFalse)
        self.Backend.IfThenElse(self, cond, ifTrue, ifFalse)

    @staticmethod
    def ifThenElse(cond: "Tensor", ifTrue: "Tensor", ifFalse: "Tensor") -> "Tensor":
        """Tensor.ifThenElse, Tensor.fs:2056-2083."""
        trgt, cond, ifTrue, ifFalse = Tensor.PrepareElemwise(ifTrue.DataType, cond, ifTrue, ifFalse)
        trgt.Backend.IfThenElse(trgt, cond, ifTrue, ifFalse)
        return trgt

    # ---- gather / scatter (Tensor.fs:2090-2198) ---------------------------------------------------------------
    def FillGather(self, indices: List[Optional["Tensor"]], src: "Tensor") -> None:
        Tensor.CheckSameStorage(src, *[i for i in indices if i is not None])
        if src.NDims != len(indices):
            raise ValueError("For each dimension of src an index tensor must be specified.")
        if any(i is None for i in indices[self.NDims:]):
            raise ValueError("Index dimensions beyond the number of target dimensions must not be None.")
        idx = [None if i is None else i.broadcastTo(self.Shape) for i in indices]
        self.Backend.Gather(self, idx, src)

    @staticmethod
    def gather(indices: List[Optional["Tensor"]], src: "Tensor") -> "Tensor":
        spec = [i for i in indices if i is not None]
        if not spec:
            raise ValueError("At least one index tensor must not be None.")
        bc = iter(Tensor.broadcastToSame(*spec))
        bc_indices = [None if i is None else next(bc) for i in indices]
        shape = next(i for i in bc_indices if i is not None).Shape
        trgt = Tensor.empty(shape, src.DataType, src.Dev)
        trgt.FillGather(bc_indices, src)
        return trgt

    def FillScatter(self, indices: List[Optional["Tensor"]], src: "Tensor") -> None:
        Tensor.CheckSameStorage(src, *[i for i in indices if i is not None])
        if self.NDims != len(indices):
            raise ValueError("For each dimension of the target an index tensor must be specified.")
        if any(i is None for i in indices[src.NDims:]):
            raise ValueError("Index dimensions beyond the number of source dimensions must not be None.")
        idx = [None if i is None else i.broadcastTo(src.Shape) for i in indices]
        # the reference zero-fills here AND in the CUDA backend (Tensor.fs:2159, CudaBackend.fs:379); the
        # backend call is self-contained, so the frontend fill is dropped.
        self.Backend.Scatter(self, idx, src)

    @staticmethod
    def scatter(indices, trgtShp, src: "Tensor") -> "Tensor":
        trgt = Tensor.empty(trgtShp, src.DataType, src.Dev)
        trgt.FillScatter(indices, src)
        return trgt

    # ---- reductions (Tensor.fs:2201-2700) -------------------------------------------------------------------
    def _fill_axis(self, member: str, ax: int, src: "Tensor", on_src_backend: bool = False) -> None:
        src = Tensor.PrepareAxisReduceSources(self, ax, src)
        getattr((src if on_src_backend else self).Backend, member)(self, src)

    def _axis(self, member: str, ax: int, dtype: Optional[int] = None) -> "Tensor":
        trgt = Tensor.PrepareAxisReduceTarget(self.DataType if dtype is None else dtype, ax, self)
        trgt._fill_axis(member, ax, self, on_src_backend=dtype is not None)
        return trgt

    def _whole(self, member: str, dtype: Optional[int] = None) -> "Tensor":
        """Tensor.sumTensor etc. (Tensor.fs:2293-2299): flatten, then reduce axis 0 into a rank-0 tensor."""
        return self.flatten()._axis(member, 0, dtype)

    def sumAxis(self, ax): return self._axis("SumLastAxis", ax)
    def productAxis(self, ax): return self._axis("ProductLastAxis", ax)
    def minAxis(self, ax): return self._axis("MinLastAxis", ax)
    def maxAxis(self, ax): return self._axis("MaxLastAxis", ax)
    def allAxis(self, ax): return self._axis("AllLastAxis", ax)
    def anyAxis(self, ax): return self._axis("AnyLastAxis", ax)
    def countTrueAxis(self, ax): return self._axis("CountTrueLastAxis", ax, dtypes.DN_I64)
    def argMinAxis(self, ax): return self._axis("ArgMinLastAxis", ax, dtypes.DN_I64)
    def argMaxAxis(self, ax): return self._axis("ArgMaxLastAxis", ax, dtypes.DN_I64)
    def sumTensor(self): return self._whole("SumLastAxis")
    def productTensor(self): return self._whole("ProductLastAxis")
    def minTensor(self): return self._whole("MinLastAxis")
    def maxTensor(self): return self._whole("MaxLastAxis")
    def allTensor(self): return self._whole("AllLastAxis")
    def anyTensor(self): return self._whole("AnyLastAxis")
    def sum(self): return self.sumTensor().Value
    def product(self): return self.productTensor().Value
    def min(self): return self.minTensor().Value
    def max(self): return self.maxTensor().Value
    def all(self): return bool(self.allTensor().Value)
    def any(self): return bool(self.anyTensor().Value)

    def countTrue(self) -> int:
        """Tensor.countTrue, Tensor.fs:2232-2233 — blocking read-back."""
        return self.Backend.CountTrue(self)

    def findAxis(self, value, ax: int) -> "Tensor":
        """Tensor.findAxis, Tensor.fs:2541-2560."""
        trgt = Tensor.PrepareAxisReduceTarget(dtypes.DN_I64, ax, self)
        src = Tensor.PrepareAxisReduceSources(trgt, ax, self)
        src.Backend.FindLastAxis(value, trgt, src)
        return trgt

    def argMax(self) -> Tuple[int, ...]:
        """Tensor.argMax, Tensor.fs:2512-2533: linear arg-reduce over the flattened tensor, then linearToIdx."""
        lin = int(self.flatten().argMaxAxis(0).Value)
        return self._linear_to_idx(lin)

    def argMin(self) -> Tuple[int, ...]:
        lin = int(self.flatten().argMinAxis(0).Value)
        return self._linear_to_idx(lin)

    def tryFind(self, value) -> Optional[Tuple[int, ...]]:
        """Tensor.tryFind, Tensor.fs:2558-2580."""
        lin = int(self.flatten().findAxis(value, 0).Value)
        return None if lin == NotFound else self._linear_to_idx(lin)

    def _linear_to_idx(self, lin: int) -> Tuple[int, ...]:
        if lin == NotFound:
            raise RuntimeError("value not found")
        idx = []
        for s in TL.cStride(self.Shape):
            idx.append(lin // s)
            lin %= s
        return tuple(idx)

    def trueIdx(self) -> "Tensor":
        """Tensor.trueIdx, Tensor.fs:2259-2263."""
        n_true = self.countTrue()
        trgt = Tensor.empty((n_true, self.NDims), dtypes.DN_I64, self.Dev)
        trgt.Backend.TrueIndices(trgt, self)
        return trgt

    # ---- masking (Tensor.fs:3011-3069) ------------------------------------------------------------------------
    @staticmethod
    def MaskShapes(masks: List[Optional["Tensor"]], shape: Tuple[int, ...]) -> List[Tuple[int, int]]:
        out = []
        shape = tuple(shape)
        for m in masks:
            if m is None:
                if not shape:
                    raise ValueError("Dimension mismatch between masks and tensor shape.")
                out.append((shape[0], shape[0]))
                shape = shape[1:]
            else:
                if m.NDims > len(shape):
                    raise ValueError("Dimension mismatch between masks and tensor shape.")
                s, shape = shape[:m.NDims], shape[m.NDims:]
                if m.Shape != s:
                    raise ValueError(f"Shape of mask {m.Shape} does not match part {s} of tensor shape it applies to.")
                out.append((m.countTrue(), m.NElems))
        if shape:
            raise ValueError("Dimension mismatch between masks and tensor shape.")
        return out

    def M(self, *masks: Optional["Tensor"]) -> "Tensor":
        """MaskedGet, Tensor.fs:3044-3052."""
        masks = list(masks)
        Tensor.CheckSameStorage(self, *[m for m in masks if m is not None])
        shapes = Tensor.MaskShapes(masks, self.Shape)
        trgt = Tensor.empty([t for t, _ in shapes], self.DataType, self.Dev)
        src = self.reshape([s for _, s in shapes])
        flat = [None if m is None else m.flatten() for m in masks]
        self.Backend.MaskedGet(trgt, src, flat)
        return trgt

    def SetM(self, masks: List[Optional["Tensor"]], value: "Tensor") -> None:
        """MaskedSet, Tensor.fs:3056-3069."""
        Tensor.CheckSameStorage(self, value, *[m for m in masks if m is not None])
        shapes = Tensor.MaskShapes(masks, self.Shape)
        value_shp, trgt_shp = [v for v, _ in shapes], [t for _, t in shapes]
        flat = [None if m is None else m.flatten() for m in masks]
        value = value.broadcastTo(value_shp)
        view = self.tryReshapeView(trgt_shp)
        if view is not None:
            self.Backend.MaskedSet(view, flat, value)
        else:
            trgt = self.reshape(trgt_shp)
            self.Backend.MaskedSet(trgt, flat, value)
            self.CopyFrom(trgt.reshape(self.Shape))

    # ---- dot (Tensor.fs:2714-2798) ----------------------------------------------------------------------------
    def FillDot(self, a: "Tensor", b: "Tensor") -> None:
        Tensor.CheckSameStorage(self, a, b)
        nd = (self.NDims, a.NDims, b.NDims)
        if nd == (0, 1, 1) and a.Shape == b.Shape:
            self.Backend.VecVecDot(self, a, b)
        elif nd == (1, 2, 1) and self.Shape[0] == a.Shape[0] and a.Shape[1] == b.Shape[0]:
            self.Backend.MatVecDot(self, a, b)
        elif nd == (2, 2, 2) and self.Shape == (a.Shape[0], b.Shape[1]) and a.Shape[1] == b.Shape[0]:
            self.Backend.MatMatDot(self, a, b)
        elif a.NDims == b.NDims and a.NDims > 2 and a.Shape[-1] == b.Shape[-2]:
            ba, bb = a.Shape[:-2], b.Shape[:-2]
            bc_a, bc_b = Tensor.broadcastToSame(a.Relayout(TL.TensorLayout(ba, 0, a.Stride[:-2])),
                                                b.Relayout(TL.TensorLayout(bb, 0, b.Stride[:-2])))
            a = a.Relayout(TensorLayout(bc_a.Shape + a.Shape[-2:], a.Offset, bc_a.Stride + a.Stride[-2:]))
            b = b.Relayout(TensorLayout(bc_b.Shape + b.Shape[-2:], b.Offset, bc_b.Stride + b.Stride[-2:]))
            if self.Shape != a.Shape[:-1] + (b.Shape[-1],):
                raise ValueError(f"Cannot compute dot product between tensors of shapes {a.Shape} and {b.Shape} "
                                 f"into tensor of shape {self.Shape}.")
            self.Backend.BatchedMatMatDot(self, a, b)
        else:
            raise ValueError(f"Cannot compute dot product between tensors of shapes {a.Shape} and {b.Shape} "
                             f"into tensor of shape {self.Shape}.")

    # ---- fused element-wise expressions (new) ------------------------------------------------
    def FillFused(self, fn, *srcs: "Tensor") -> None:
        """Evaluates `fn(*srcs)` — an expression of +, -, *, /, %, **, unary minus, abs and the unary functions
        (sin, exp, tanh, ...) over up to three tensors and Python scalars — in ONE backend call
        (ITensorBackend extension `FusedElemwise`), with the rounding of the operator-by-operator evaluation."""
        from .fused import trace
        srcs = Tensor.PrepareElemwiseSources(self, *srcs)
        prog = trace(fn, len(srcs))
        self.Backend.FusedElemwise(self, list(srcs), prog)

    @staticmethod
    def fused(fn, *srcs: "Tensor") -> "Tensor":
        trgt, *bsrcs = Tensor.PrepareElemwise(srcs[0].DataType, *srcs)
        from .fused import trace
        trgt.Backend.FusedElemwise(trgt, list(bsrcs), trace(fn, len(bsrcs)))
        return trgt

    def FillInvert(self, a: "Tensor") -> None:
        """Tensor.FillInvert, Tensor.fs:2809-2815."""
        Tensor.CheckSameStorage(self, a)
        if a.NDims < 2:
            raise ValueError(f"Need at least a matrix to invert but got shape {a.Shape}.")
        a = a.broadcastTo(self.Shape)
        self.Backend.BatchedInvert(self, a)

    @staticmethod
    def invert(a: "Tensor") -> "Tensor":
        """Tensor.invert, Tensor.fs:2836-2839: (batch) inverse of [..., n, n]; SingularMatrixException if singular."""
        trgt = Tensor.empty(a.Shape, a.DataType, a.Dev)
        trgt.FillInvert(a)
        return trgt

    def __matmul__(self, b: "Tensor") -> "Tensor":
        """(.*), Tensor.fs:2772-2798."""
        a = self
        if a.NDims == 1 and b.NDims == 1:
            shp = ()
        elif a.NDims == 2 and b.NDims == 1:
            shp = (a.Shape[0],)
        elif a.NDims == 2 and b.NDims == 2:
            shp = (a.Shape[0], b.Shape[1])
        elif a.NDims == b.NDims and a.NDims > 2:
            batch = TL.broadcastToSameMany([TL.newC(a.Shape[:-2]), TL.newC(b.Shape[:-2])])[0].Shape
            shp = batch + (a.Shape[-2], b.Shape[-1])
        else:
            raise ValueError(f"Cannot compute dot product between tensors of shapes {a.Shape} and {b.Shape}.")
        trgt = Tensor.empty(shp, a.DataType, a.Dev)
        trgt.FillDot(a, b)
        return trgt

    def __repr__(self):
        return f"Tensor<{dtypes.NAMES[self.DataType]}>(shape={self.Shape}, stride={self.Stride}, " \
               f"offset={self.Offset}, dev={self.Dev})"


def _install_unary_functions():
    """sgn, log, log10, exp, sin … truncate as methods + Fill* variants (Tensor.fs:836-1300)."""
    names = {"Abs": "abs", "Sgn": "sgn", "Log": "log", "Log10": "log10", "Exp": "exp", "Sin": "sin", "Cos": "cos",
             "Tan": "tan", "Asin": "asin", "Acos": "acos", "Atan": "atan", "Sinh": "sinh", "Cosh": "cosh",
             "Tanh": "tanh", "Sqrt": "sqrt", "Ceiling": "ceil", "Floor": "floor", "Round": "round",
             "Truncate": "truncate"}
    for member, fn in names.items():
        setattr(Tensor, fn, (lambda m: lambda self: self._new_unary(m))(member))
    for member in ["UnaryPlus", "UnaryMinus", "Negate"] + list(names):
        setattr(Tensor, "Fill" + member, (lambda m: lambda self, a: self._fill_unary(m, a))(member))
    for member in ["Add", "Subtract", "Multiply", "Divide", "Modulo", "Power", "MaxElemwise", "MinElemwise",
                   "And", "Or", "Xor", "Equal", "NotEqual", "Less", "LessOrEqual", "Greater", "GreaterOrEqual"]:
        setattr(Tensor, "Fill" + member, (lambda m: lambda self, a, b: self._fill_binary(m, a, b))(member))
    for member, fn in [("SumLastAxis", "FillSumAxis"), ("ProductLastAxis", "FillProductAxis"),
                       ("MinLastAxis", "FillMinAxis"), ("MaxLastAxis", "FillMaxAxis"),
                       ("AllLastAxis", "FillAllAxis"), ("AnyLastAxis", "FillAnyAxis")]:
        setattr(Tensor, fn, (lambda m: lambda self, ax, src: self._fill_axis(m, ax, src))(member))


_install_unary_functions()


class CudaTensor:
    """module CudaTensor, Tensor/Tensor/Cuda/CudaFrontend.fs:34-150."""
    Dev: TensorCudaDevice = None  # set lazily: constructing it loads libdeepnet_b200.so

    @staticmethod
    def dev() -> TensorCudaDevice:
        if CudaTensor.Dev is None:
            CudaTensor.Dev = TensorCudaDevice.Instance()
        return CudaTensor.Dev

    @staticmethod
    def transfer(x: Tensor) -> Tensor:
        return Tensor.transfer(CudaTensor.dev(), x)

    @staticmethod
    def ofNumpy(arr: np.ndarray) -> Tensor:
        return Tensor.ofNumpy(arr, CudaTensor.dev())

    @staticmethod
    def zeros(shape, dtype) -> Tensor:
        return Tensor.zeros(shape, dtype, CudaTensor.dev())

    @staticmethod
    def usingPtr(ptr: int, shape, dtype: int, owner=None) -> Tensor:
        """CudaTensor.usingPtr, CudaFrontend.fs:129-137: wrap external device memory without owning it."""
        lay = TL.newC(shape)
        return Tensor(lay, CudaTensor.dev().UsingPtr(ptr, lay.NElems, dtype, owner))
