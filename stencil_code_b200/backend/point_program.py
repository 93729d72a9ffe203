"""Typed per-point program: what is left of a stencil kernel after neighbor unrolling.

The reference never names this stage -- ``StencilBackend.visit_NeighborPointsLoop``
(``backend/stencil_backend.py:67-90``) splices the unrolled statements straight into a ctree C
loop nest and lets gcc apply C's typing rules.  The CUDA backend needs the same statements as
data (to plan tiles from the tap footprint, to classify arguments, and to prove when a double
expression may run in float32 without changing a bit), so they are kept here as a small typed
expression IR.  Types are the C types the reference's generated code would have:

  * Python ``int`` literal -> C ``int``; Python ``float`` literal -> C ``double``
    (ctree prints ``str(value)``; SURVEY.md section 7 "literal typing")
  * grid / table element -> the array's element type
  * binary operators     -> C's usual arithmetic conversions
  * ``int(e)``           -> ``(int)e`` (truncation), ``abs`` -> ``int abs(int)``
    (reference ``backend/c.py:22-24``), ``min`` -> the reference's ternary macro (``:25-27``)
"""
import math

import numpy

from ..stencil_exception import StencilException

INT, LONG, FLOAT, DOUBLE = 'int', 'long long', 'float', 'double'
_RANK = {INT: 0, LONG: 1, FLOAT: 2, DOUBLE: 3}

_DTYPE_TO_CTYPE = {
    numpy.dtype('float32'): FLOAT, numpy.dtype('float64'): DOUBLE,
    numpy.dtype('int32'): INT, numpy.dtype('int64'): LONG,
}


def ctype_of_dtype(dtype):
    try:
        return _DTYPE_TO_CTYPE[numpy.dtype(dtype)]
    except KeyError:
        raise StencilException("unsupported array dtype {} (float32, float64, int32, int64 "
                               "are supported)".format(dtype))


def common_type(a, b):
    """C usual arithmetic conversions restricted to the four types above."""
    return a if _RANK[a] >= _RANK[b] else b


class Expr(object):
    ctype = None

    def children(self):
        return ()


class Const(Expr):
    def __init__(self, value, ctype=None):
        if ctype is None:
            if isinstance(value, (bool, numpy.bool_)):
                value, ctype = int(value), INT
            elif isinstance(value, (int, numpy.integer)):
                value, ctype = int(value), INT
            elif isinstance(value, (float, numpy.floating)):
                value, ctype = float(value), DOUBLE
            else:
                raise StencilException("constant {!r} is neither int nor float".format(value))
        self.value = value
        self.ctype = ctype

    def c_text(self):
        if self.ctype in (INT, LONG):
            return str(int(self.value))
        if not math.isfinite(self.value):
            raise StencilException("non-finite literal {!r} cannot be specialized".format(self.value))
        if self.ctype == FLOAT:
            text = repr(float(numpy.float32(self.value)))
            return text + 'f'
        return repr(float(self.value))


class Tap(Expr):
    """Element of grid argument ``arg`` at (current point + offset)."""
    def __init__(self, arg, offset, ctype):
        self.arg, self.offset, self.ctype = arg, tuple(int(o) for o in offset), ctype


class TableRef(Expr):
    """Element of 1-d table argument ``arg`` at a computed index."""
    def __init__(self, arg, index, ctype):
        self.arg, self.index, self.ctype = arg, index, ctype

    def children(self):
        return (self.index,)


class OutRef(Expr):
    """The output element of the current point (the accumulator), read."""
    def __init__(self, ctype):
        self.ctype = ctype


class LocalRef(Expr):
    def __init__(self, name, ctype):
        self.name, self.ctype = name, ctype


class Bin(Expr):
    def __init__(self, op, left, right, ctype=None):
        self.op, self.left, self.right = op, left, right
        self.ctype = ctype or common_type(left.ctype, right.ctype)

    def children(self):
        return (self.left, self.right)


class Neg(Expr):
    def __init__(self, operand):
        self.operand, self.ctype = operand, operand.ctype

    def children(self):
        return (self.operand,)


class Cast(Expr):
    def __init__(self, ctype, operand):
        self.ctype, self.operand = ctype, operand

    def children(self):
        return (self.operand,)


class Call(Expr):
    """``abs`` (int abs(int)) or ``min`` (ternary macro)."""
    def __init__(self, func, args):
        self.func, self.args = func, list(args)
        if func == 'abs':
            if len(args) != 1:
                raise StencilException("abs() takes one argument")
            self.ctype = INT
        elif func == 'min':
            if len(args) != 2:
                raise StencilException("min() takes two arguments")
            self.ctype = common_type(args[0].ctype, args[1].ctype)
        else:
            raise StencilException(
                "call to '{}' cannot be specialized (abs, min, int and self.distance are "
                "available, as in the reference C prologue)".format(func))

    def children(self):
        return tuple(self.args)


class Store(object):
    """``out[p] op expr`` with op in '=', '+=', '-=', '*=', '/='."""
    def __init__(self, op, expr):
        self.op, self.expr = op, expr


class LocalAssign(object):
    """``name op expr`` for a scalar temporary (an additive extension; the reference has none)."""
    def __init__(self, name, op, expr, ctype, declares):
        self.name, self.op, self.expr, self.ctype, self.declares = name, op, expr, ctype, declares


class ArgInfo(object):
    """One kernel argument as the backend sees it."""
    def __init__(self, name, index, shape, dtype):
        self.name, self.index = name, index
        self.shape, self.dtype = tuple(int(s) for s in shape), numpy.dtype(dtype)
        self.ctype = ctype_of_dtype(dtype)
        self.used_as_grid = False     # indexed by the point / neighbor target
        self.used_as_table = False    # indexed by a computed expression

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


class PointProgram(object):
    """Everything a code generator needs to know about one specialised stencil."""
    def __init__(self, dim, shape, out_ctype, args, statements, ghost_depth, boundary):
        self.dim = dim
        self.shape = tuple(int(s) for s in shape)
        self.out_ctype = out_ctype
        self.args = args                  # list[ArgInfo], inputs only (output excluded)
        self.statements = statements      # list[Store | LocalAssign]
        self.ghost_depth = tuple(int(g) for g in ghost_depth)
        self.boundary = boundary          # 'clamp' | 'zero' | 'copy' | 'periodic'   ('wrap' arrives as 'zero')

    @property
    def out_dtype_size(self):
        return 8 if self.out_ctype == DOUBLE else 4

    @property
    def npoints(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def walk(self):
        """All expression nodes of all statements, pre-order."""
        stack = [s.expr for s in reversed(self.statements)]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children()))

    def taps(self):
        return [n for n in self.walk() if isinstance(n, Tap)]

    def tap_extent(self):
        """Per-dimension (most negative, most positive) tap offset, origin included."""
        lo = [0] * self.dim
        hi = [0] * self.dim
        for tap in self.taps():
            for d, o in enumerate(tap.offset):
                lo[d] = min(lo[d], o)
                hi[d] = max(hi[d], o)
        return tuple(lo), tuple(hi)

    def uses_double(self):
        return self.out_ctype == DOUBLE or any(n.ctype == DOUBLE for n in self.walk())


# ---- C text -----------------------------------------------------------------------------------
def expr_to_c(expr, tap_text, table_text, out_text='acc'):
    """C/CUDA text of an expression.  Every binary node is parenthesised, so the tree shape --
    and with it the evaluation order the reference's C would have -- is explicit."""
    def emit(e):
        if isinstance(e, Const):
            text = e.c_text()
            return '(' + text + ')' if text.startswith('-') else text
        if isinstance(e, Tap):
            return tap_text(e)
        if isinstance(e, TableRef):
            return table_text(e, emit(e.index))
        if isinstance(e, OutRef):
            return out_text
        if isinstance(e, LocalRef):
            return e.name
        if isinstance(e, Bin):
            return '({} {} {})'.format(emit(e.left), e.op, emit(e.right))
        if isinstance(e, Neg):
            return '(-{})'.format(emit(e.operand))
        if isinstance(e, Cast):
            return '(({}){})'.format(e.ctype, emit(e.operand))
        if isinstance(e, Call):
            if e.func == 'abs':
                return 'abs((int){})'.format(emit(e.args[0]))
            a, b = emit(e.args[0]), emit(e.args[1])
            return '(({a} < {b}) ? {a} : {b})'.format(a=a, b=b)
        raise StencilException("cannot emit {!r}".format(e))
    return emit(expr)


# ---- exact float32 demotion -------------------------------------------------------------------
def _is_float_exact_const(value):
    return math.isfinite(value) and float(numpy.float32(value)) == float(value)


def _is_pow2(value):
    if value == 0 or not math.isfinite(value):
        return False
    mantissa, _ = math.frexp(abs(value))
    return mantissa == 0.5


def _exact_in_float(expr):
    """If the double-typed ``expr`` always holds a value that float32 represents exactly and
    that float32 arithmetic produces without rounding, return that float expression; else None.

    * a float-typed sub-expression promoted to double: itself
    * a double literal that is exactly a float32: the float literal
    * (+-2^k) * x with x as above: the product is exact in binary floating point as long as it
      neither overflows float32 nor falls into its subnormal range.  Precondition (DESIGN.md
      section 4): |x| * 2^k within [2^-126, 2^128) -- any data of physical magnitude; inputs near
      FLT_MAX or FLT_MIN need STENCIL_CUDA_STRICT=1, which keeps the double arithmetic
      (tests/test_gpu_parity.py::test_extreme_magnitudes_need_strict_mode).  1.0 * x is x.
    """
    if expr.ctype in (INT, LONG):
        return None
    if expr.ctype == FLOAT:
        return expr
    if isinstance(expr, Const):
        return Const(expr.value, FLOAT) if _is_float_exact_const(expr.value) else None
    if isinstance(expr, Cast) and expr.ctype == DOUBLE:
        return _exact_in_float(expr.operand)
    if isinstance(expr, Bin) and expr.op == '*':
        for const, other in ((expr.left, expr.right), (expr.right, expr.left)):
            if isinstance(const, Const) and const.ctype == DOUBLE and _is_pow2(const.value) \
                    and _is_float_exact_const(const.value):
                inner = _exact_in_float(other)
                if inner is not None:
                    if const.value == 1.0:
                        return inner                  # 1.0 * x is x, bit for bit (-0.0, inf, NaN included)
                    if const is expr.left:
                        return Bin('*', Const(const.value, FLOAT), inner, FLOAT)
                    return Bin('*', inner, Const(const.value, FLOAT), FLOAT)
    return None


def _single_op_in_float(expr):
    """``(float)(a op b)`` computed in double equals ``a op b`` computed in float32 when a and b
    are exact float32 values: double's 53 bits >= 2*24+2, so the double rounding is innocuous
    for + - * / (Figueroa 1995).  Returns the float expression or None."""
    if isinstance(expr, Bin) and expr.ctype == DOUBLE and expr.op in '+-*/':
        left, right = _exact_in_float(expr.left), _exact_in_float(expr.right)
        if left is not None and right is not None:
            if expr.op == '*':                        # 1.0 * x is x, bit for bit
                for const, other in ((left, right), (right, left)):
                    if isinstance(const, Const) and const.value == 1.0:
                        return other
            return Bin(expr.op, left, right, FLOAT)
    return _exact_in_float(expr)


def demote_exact(program):
    """Rewrite double sub-computations as float32 where the stored float32 result is provably
    bit-identical (see ``_exact_in_float`` / ``_single_op_in_float``).  Only touches stores to a
    float32 output; returns the number of statements rewritten."""
    if program.out_ctype != FLOAT:
        return 0
    rewritten = 0
    for statement in program.statements:
        if not isinstance(statement, Store) or statement.expr.ctype != DOUBLE:
            continue
        if statement.op == '=':
            replacement = _single_op_in_float(statement.expr)
        else:
            # acc op= E  is  (float)((double)acc op E): one operation on two exact floats
            replacement = _exact_in_float(statement.expr)
        if replacement is not None:
            statement.expr = replacement
            rewritten += 1
    return rewritten
