"""TensorLayout — shape / offset / stride views, mirroring Tensor/Tensor/TensorLayout.fs of the reference.

Only layout arithmetic lives here (no data): the frontend builds views by transforming layouts and the backend
receives them as `dn_tensor` descriptors. Function names follow the reference module (`TensorLayout.swapDim`,
`broadcastDim`, `reverseAxis`, `permuteAxes`, `tryReshape`, `view`, `diagAxis`), file:line cited per function.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

# Tensor/Tensor/TensorRng.fs:11-24
NewAxis = -(2 ** 63) + 1
Fill = -(2 ** 63) + 2
Remainder = -(2 ** 63) + 3
NotFound = -(2 ** 63) + 4


def _prod(xs: Sequence[int]) -> int:
    n = 1
    for x in xs:
        n *= x
    return n


@dataclass(frozen=True)
class TensorLayout:
    """TensorLayout.fs:12-24."""
    Shape: Tuple[int, ...]
    Offset: int
    Stride: Tuple[int, ...]

    def __post_init__(self):
        # TensorLayout.check, TensorLayout.fs:31-36
        if len(self.Shape) != len(self.Stride):
            raise ValueError(f"shape {self.Shape} and stride {self.Stride} must have same number of entries")
        if any(s < 0 for s in self.Shape):
            raise ValueError(f"shape {self.Shape} cannot have negative entries")

    @property
    def NDims(self) -> int:
        return len(self.Shape)

    @property
    def NElems(self) -> int:
        return _prod(self.Shape)


def cStride(shape: Sequence[int]) -> Tuple[int, ...]:
    """Row-major strides, TensorLayout.fs:107-108."""
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def fStride(shape: Sequence[int]) -> Tuple[int, ...]:
    """Column-major strides, TensorLayout.fs:111-112."""
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= s
    return tuple(st)


def newC(shape: Sequence[int]) -> TensorLayout:
    return TensorLayout(tuple(shape), 0, cStride(shape))


def newF(shape: Sequence[int]) -> TensorLayout:
    return TensorLayout(tuple(shape), 0, fStride(shape))


def stridesEqual(shape, a, b) -> bool:
    """TensorLayout.fs:131-133."""
    return all(x == y for s, x, y in zip(shape, a, b) if s > 1)


def isC(a: TensorLayout) -> bool:
    return stridesEqual(a.Shape, a.Stride, cStride(a.Shape))


def isF(a: TensorLayout) -> bool:
    return stridesEqual(a.Shape, a.Stride, fStride(a.Shape))


def hasContiguousMemory(a: TensorLayout) -> bool:
    """TensorLayout.fs:144-146."""
    return isC(a) or isF(a)


def checkAxis(ax: int, a: TensorLayout) -> None:
    if not (0 <= ax < a.NDims):
        raise IndexError(f"axis {ax} out of range for tensor with shape {a.Shape}")


def addr(idx: Sequence[int], a: TensorLayout) -> int:
    """TensorLayout.fs:46-48."""
    if len(idx) != a.NDims or not all(0 <= i < s for i, s in zip(idx, a.Shape)):
        raise IndexError(f"index {tuple(idx)} out of range for tensor of shape {a.Shape}")
    return a.Offset + sum(i * s for i, s in zip(idx, a.Stride))


def padLeft(a: TensorLayout) -> TensorLayout:
    return TensorLayout((1,) + a.Shape, a.Offset, (0,) + a.Stride)


def padRight(a: TensorLayout) -> TensorLayout:
    return TensorLayout(a.Shape + (1,), a.Offset, a.Stride + (0,))


def insertAxis(ax: int, a: TensorLayout) -> TensorLayout:
    if not (0 <= ax <= a.NDims):
        raise IndexError(f"axis {ax} out of range for tensor with shape {a.Shape}")
    return TensorLayout(a.Shape[:ax] + (1,) + a.Shape[ax:], a.Offset, a.Stride[:ax] + (0,) + a.Stride[ax:])


def cutLeft(a: TensorLayout) -> TensorLayout:
    if a.NDims == 0:
        raise ValueError("cannot remove dimensions from scalar")
    return TensorLayout(a.Shape[1:], a.Offset, a.Stride[1:])


def cutRight(a: TensorLayout) -> TensorLayout:
    if a.NDims == 0:
        raise ValueError("cannot remove dimensions from scalar")
    return TensorLayout(a.Shape[:-1], a.Offset, a.Stride[:-1])


def broadcastDim(dim: int, size: int, a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:178-182."""
    if size < 0:
        raise ValueError("size must be positive")
    if a.Shape[dim] != 1:
        raise RuntimeError(f"Dimension {dim} of shape {a.Shape} must be of size 1 to broadcast.")
    shp, st = list(a.Shape), list(a.Stride)
    shp[dim], st[dim] = size, 0
    return TensorLayout(tuple(shp), a.Offset, tuple(st))


def padToSameMany(sas: List[TensorLayout]) -> List[TensorLayout]:
    need = max(s.NDims for s in sas)
    out = []
    for sa in sas:
        while sa.NDims < need:
            sa = padLeft(sa)
        out.append(sa)
    return out


def broadcastToSameMany(sas: List[TensorLayout]) -> List[TensorLayout]:
    """TensorLayout.fs:211-254."""
    if not sas:
        return []
    sas = padToSameMany(list(sas))
    for d in range(sas[0].NDims):
        ls = [sa.Shape[d] for sa in sas]
        if any(l == 1 for l in ls):
            non_bc = {l for l in ls if l != 1}
            if len(non_bc) == 1:
                target = next(iter(non_bc))
                sas = [broadcastDim(d, target, sa) if sa.Shape[d] != target else sa for sa in sas]
            elif len(non_bc) > 1:
                raise RuntimeError(f"Cannot broadcast shapes {[s.Shape for s in sas]} to same size.")
        elif len(set(ls)) > 1:
            raise RuntimeError(f"Cannot broadcast shapes {[s.Shape for s in sas]} to same size.")
    return sas


def broadcastToShape(bs: Sequence[int], ain: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:257-271."""
    bs = tuple(bs)
    if len(bs) < ain.NDims:
        raise RuntimeError(f"Cannot broadcast to shape {bs} from shape {ain.Shape} of higher rank.")
    a = ain
    while a.NDims < len(bs):
        a = padLeft(a)
    for d in range(len(bs)):
        if a.Shape[d] == bs[d]:
            continue
        if a.Shape[d] == 1:
            a = broadcastDim(d, bs[d], a)
        else:
            raise RuntimeError(f"Cannot broadcast shape {ain.Shape} to shape {bs}.")
    return a


def isBroadcasted(a: TensorLayout) -> bool:
    return any(st == 0 and sh > 1 for sh, st in zip(a.Shape, a.Stride))


def _resolve_remainder(shp: Sequence[int], nelems: int) -> Tuple[int, ...]:
    shp = list(shp)
    n_rem = sum(1 for s in shp if s == Remainder)
    if n_rem == 0:
        return tuple(shp)
    if n_rem > 1:
        raise ValueError(f"only the size of one dimension can be determined automatically, but shape was {shp}")
    so_far = _prod([s for s in shp if s != Remainder])
    if so_far == 0 or nelems % so_far != 0:
        raise ValueError(f"cannot reshape to {shp}: {nelems} / {so_far} is not an integer")
    return tuple(nelems // so_far if s == Remainder else s for s in shp)


def tryReshape(shp: Sequence[int], a: TensorLayout) -> Optional[TensorLayout]:
    """TensorLayout.fs:282-325: reshape without copying, or None if a copy is required."""
    shp = _resolve_remainder(shp, a.NElems)
    if _prod(shp) != a.NElems:
        raise ValueError(f"cannot reshape from shape {a.Shape} ({a.NElems} elements) to shape {shp}")
    if isC(a):
        return TensorLayout(shp, a.Offset, cStride(shp))

    def tf(new_str, new_shp, a_str, a_shp):
        if new_shp and a_str and a_shp and new_shp[0] == a_shp[0]:
            return tf(new_str + [a_str[0]], new_shp[1:], a_str[1:], a_shp[1:])
        if new_shp and new_shp[0] == 1:
            return tf(new_str + [0], new_shp[1:], a_str, a_shp)
        if a_str and a_shp and a_shp[0] == 1:
            return tf(new_str, new_shp, a_str[1:], a_shp[1:])
        if not new_shp and not a_str and not a_shp:
            return new_str
        return None

    st = tf([], list(shp), list(a.Stride), list(a.Shape))
    return None if st is None else TensorLayout(shp, a.Offset, tuple(st))


def swapDim(ax1: int, ax2: int, a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:344-349."""
    if not (0 <= ax1 < a.NDims and 0 <= ax2 < a.NDims):
        raise ValueError(f"Cannot swap dimension {ax1} with {ax2} for shape {a.Shape}.")
    shp, st = list(a.Shape), list(a.Stride)
    shp[ax1], shp[ax2] = shp[ax2], shp[ax1]
    st[ax1], st[ax2] = st[ax2], st[ax1]
    return TensorLayout(tuple(shp), a.Offset, tuple(st))


def transpose(a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:353-356: swaps the last two axes."""
    if a.NDims < 2:
        raise ValueError(f"cannot transpose non-matrix of shape {a.Shape}")
    return swapDim(a.NDims - 2, a.NDims - 1, a)


def permuteAxes(permut: Sequence[int], a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:361-365: permut[i] is the NEW position of axis i."""
    if len(permut) != a.NDims or sorted(permut) != list(range(a.NDims)):
        raise ValueError(f"Permutation {permut} must have same rank as shape {a.Shape}.")
    shp, st = [0] * a.NDims, [0] * a.NDims
    for i, p in enumerate(permut):
        shp[p], st[p] = a.Shape[i], a.Stride[i]
    return TensorLayout(tuple(shp), a.Offset, tuple(st))


def reverseAxis(ax: int, a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:368-371."""
    checkAxis(ax, a)
    st = list(a.Stride)
    st[ax] = -st[ax]
    return TensorLayout(a.Shape, a.Offset + (a.Shape[ax] - 1) * a.Stride[ax], tuple(st))


Rng = Union[int, slice, None, type(Ellipsis)]


def view(ranges: Sequence[Rng], a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:374-427. Ranges: int = Rng.Elem, slice (no step; stop EXCLUSIVE as in Python, i.e. the
    reference's inclusive `last` + 1) = Rng.Rng, None = Rng.NewAxis, Ellipsis = Rng.AllFill."""
    ranges = list(ranges)
    n_consuming = sum(1 for r in ranges if r is not None and r is not Ellipsis)
    if Ellipsis in ranges:
        i = ranges.index(Ellipsis)
        ranges[i:i + 1] = [slice(None)] * (a.NDims - n_consuming)
    elif n_consuming < a.NDims:
        ranges += [slice(None)] * (a.NDims - n_consuming)
    shape, stride, offset = [], [], a.Offset
    d = 0
    for r in ranges:
        if r is None:
            shape.append(1)
            stride.append(0)
            continue
        if d >= a.NDims:
            raise IndexError(f"Slice {ranges} is incompatible with shape {a.Shape}.")
        shp, st = a.Shape[d], a.Stride[d]
        if isinstance(r, slice):
            if r.step not in (None, 1):
                raise IndexError("Deep.Net ranges have no step (TensorRng.fs:30-38)")
            start = 0 if r.start is None else (r.start + shp if r.start < 0 else r.start)
            stop = shp if r.stop is None else (r.stop + shp if r.stop < 0 else r.stop)
            if stop > start:
                if not (0 <= start < shp) or not (0 < stop <= shp):
                    raise IndexError(f"Index out of range in slice {ranges} for shape {a.Shape}.")
                offset += start * st
                shape.append(stop - start)
            else:
                shape.append(0)
            stride.append(st)
        else:
            i = int(r)
            if i < 0:
                i += shp
            if not (0 <= i < shp):
                raise IndexError(f"Index {r} out of range in slice {ranges} for shape {a.Shape}.")
            offset += i * st
        d += 1
    if d != a.NDims:
        raise IndexError(f"Slice {ranges} is incompatible with shape {a.Shape}.")
    return TensorLayout(tuple(shape), offset, tuple(stride))


def diagAxis(ax1: int, ax2: int, a: TensorLayout) -> TensorLayout:
    """TensorLayout.fs:446-461."""
    checkAxis(ax1, a)
    checkAxis(ax2, a)
    if ax1 == ax2:
        raise ValueError("Axes to use for diagonal must be different.")
    if a.Shape[ax1] != a.Shape[ax2]:
        raise ValueError(f"Array must have same dimensions along axis {ax1} and {ax2} but has shape {a.Shape}.")
    shp, st = [], []
    for ax, (sh, s) in enumerate(zip(a.Shape, a.Stride)):
        if ax == ax1:
            shp.append(sh)
            st.append(a.Stride[ax1] + a.Stride[ax2])
        elif ax != ax2:
            shp.append(sh)
            st.append(s)
    return TensorLayout(tuple(shp), a.Offset, tuple(st))
