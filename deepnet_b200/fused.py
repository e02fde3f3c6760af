"""Builds dn_fused_elemwise programs (include/dn_tensor.h) by tracing a Python expression.

    prog = trace(lambda a, b: a * b + a.sin(), 2)

The traced function receives one `Expr` per source tensor; arithmetic operators and the unary functions of the Tensor
frontend (`sin`, `exp`, `tanh`, `sqrt`, `abs`, ...) record instructions; Python numbers become CONST instructions.
Registers are allocated greedily: a value's register is released after its last use, so expression trees of any
width that fit the six virtual registers can be expressed. Common subexpressions (the same `Expr` object used
twice) are evaluated once."""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

from .backend import BINARY_OPS, UNARY_OPS
from .native import DN_FUSED_BINARY, DN_FUSED_CONST, DN_FUSED_MAX_INSTRS, DN_FUSED_REGS, DN_FUSED_UNARY


class Expr:
    def __init__(self, kind, op=0, args=(), value=0.0, src=-1):
        self.kind, self.op, self.args, self.value, self.src = kind, op, tuple(args), value, src

    @staticmethod
    def lift(x) -> "Expr":
        return x if isinstance(x, Expr) else Expr("const", value=float(x))

    def _bin(self, name, other, swap=False):
        a, b = (Expr.lift(other), self) if swap else (self, Expr.lift(other))
        return Expr("binary", BINARY_OPS.index(name), (a, b))

    def _un(self, name):
        return Expr("unary", UNARY_OPS.index(name), (self,))

    def __add__(self, o): return self._bin("Add", o)
    def __radd__(self, o): return self._bin("Add", o, True)
    def __sub__(self, o): return self._bin("Subtract", o)
    def __rsub__(self, o): return self._bin("Subtract", o, True)
    def __mul__(self, o): return self._bin("Multiply", o)
    def __rmul__(self, o): return self._bin("Multiply", o, True)
    def __truediv__(self, o): return self._bin("Divide", o)
    def __rtruediv__(self, o): return self._bin("Divide", o, True)
    def __mod__(self, o): return self._bin("Modulo", o)
    def __pow__(self, o): return self._bin("Power", o)
    def __neg__(self): return self._un("UnaryMinus")
    def __pos__(self): return self._un("UnaryPlus")
    def __abs__(self): return self._un("Abs")
    def maxElemwise(self, o): return self._bin("MaxElemwise", o)
    def minElemwise(self, o): return self._bin("MinElemwise", o)


for _member, _fn in {"Abs": "abs", "Sgn": "sgn", "Log": "log", "Log10": "log10", "Exp": "exp", "Sin": "sin",
                     "Cos": "cos", "Tan": "tan", "Asin": "asin", "Acos": "acos", "Atan": "atan", "Sinh": "sinh",
                     "Cosh": "cosh", "Tanh": "tanh", "Sqrt": "sqrt", "Ceiling": "ceil", "Floor": "floor",
                     "Round": "round", "Truncate": "truncate"}.items():
    setattr(Expr, _fn, (lambda m: lambda self: self._un(m))(_member))


_CACHE: Dict[tuple, list] = {}


def trace(fn: Callable, nsrc: int) -> List[Tuple[int, int, int, int, int, float]]:
    """Returns [(kind, op, dst, a, b, imm)] for `fn` applied to `nsrc` sources. Traces are cached per code object
    and captured constants, so a training loop pays for tracing once."""
    key = None
    code = getattr(fn, "__code__", None)
    if code is not None:
        try:
            cells = tuple(c.cell_contents for c in (fn.__closure__ or ()))
            key = (code, cells, fn.__defaults__, nsrc)
            hash(key)
        except (TypeError, ValueError):
            key = None
    if key is not None and key in _CACHE:
        return _CACHE[key]
    prog = _trace(fn, nsrc)
    if key is not None:
        if len(_CACHE) > 512:
            _CACHE.clear()
        _CACHE[key] = prog
    return prog


def _trace(fn: Callable, nsrc: int) -> List[Tuple[int, int, int, int, int, float]]:
    root = Expr.lift(fn(*[Expr("src", src=i) for i in range(nsrc)]))
    if root.kind == "src":
        root = root._un("UnaryPlus")
    # post-order over the DAG, counting uses so that registers can be released
    order: List[Expr] = []
    uses: Dict[int, int] = {}
    seen = set()

    def visit(e: Expr):
        if id(e) in seen:
            return
        seen.add(id(e))
        for a in e.args:
            uses[id(a)] = uses.get(id(a), 0) + 1
            visit(a)
        order.append(e)

    visit(root)
    free = list(range(nsrc, DN_FUSED_REGS))
    reg: Dict[int, int] = {}
    prog = []
    for e in order:
        if e.kind == "src":
            reg[id(e)] = e.src
            continue
        operands = [reg[id(a)] for a in e.args]
        # release operand registers whose last use this is (sources included: their registers are reusable too)
        for a in e.args:
            uses[id(a)] -= 1
        released = []
        for a in e.args:
            if uses[id(a)] == 0 and reg[id(a)] not in released:
                released.append(reg[id(a)])
        free = sorted(set(free) | set(released))
        if not free:
            raise ValueError("expression needs more than %d live values" % DN_FUSED_REGS)
        dst = free.pop(0)
        reg[id(e)] = dst
        if e.kind == "const":
            prog.append((DN_FUSED_CONST, 0, dst, 0, 0, e.value))
        elif e.kind == "unary":
            prog.append((DN_FUSED_UNARY, e.op, dst, operands[0], 0, 0.0))
        else:
            prog.append((DN_FUSED_BINARY, e.op, dst, operands[0], operands[1], 0.0))
    if len(prog) > DN_FUSED_MAX_INSTRS:
        raise ValueError("expression needs %d instructions, at most %d are supported" % (len(prog), DN_FUSED_MAX_INSTRS))
    return prog
