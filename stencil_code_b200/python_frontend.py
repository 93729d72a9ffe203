"""Python front end: the AST of a user's ``kernel`` method -> stencil semantic model.

Does the job of reference ``stencil_code/python_frontend.py:13-84`` (``PythonToStencilModel``)
on the standard library ``ast`` of Python >= 3.9 (the reference's ``node.slice.value`` at
``:78`` and its ``is 'interior_points'`` identity tests at ``:46,49`` no longer work there) and
without ctree's ``PyBasicConversions``.

Recognised, as in the reference:
  * ``for p in self.interior_points(grid)``  -> InteriorPointsLoop (a ``stride=`` argument is
    accepted and ignored, reference ``:46-48``)
  * ``for n in self.neighbors(p, k)``        -> NeighborPointsLoop (k defaults to 0, ``:51``)
  * ``grid[index]``                          -> GridElement (``:76-84``)
  * ``self.distance(a, b)`` / ``int(e)``     -> MathFunction (``:59-74``)
  * numbers, names, + - * /, unary minus, ``=`` and augmented assignment, plain calls
"""
import ast
import inspect
import textwrap

from .stencil_exception import StencilException
from .stencil_model import (
    KernelFunction, InteriorPointsLoop, NeighborPointsLoop, GridElement, MathFunction,
    Constant, SymbolRef, SelfAttribute, BinaryOp, UnaryOp, FunctionCall, Assign, AugAssign)

_BIN_OPS = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}
_UNARY_OPS = {ast.USub: '-', ast.UAdd: '+'}


def get_ast(function):
    """Python AST (a Module holding one FunctionDef) of a function or bound method."""
    source = textwrap.dedent(inspect.getsource(function))
    return ast.parse(source)


class PythonToStencilModel(ast.NodeVisitor):
    def __init__(self, arg_names=None):
        self.arg_names = arg_names

    # ---- structure ------------------------------------------------------------------------
    def visit_Module(self, node):
        functions = [n for n in node.body if isinstance(n, ast.FunctionDef)]
        if len(functions) != 1:
            raise StencilException("kernel source must hold exactly one function definition")
        return self.visit(functions[0])

    def visit_FunctionDef(self, node):
        params = [a.arg for a in node.args.args]
        if params and params[0] == 'self':          # strip self (reference :20-22)
            params = params[1:]
        body = []
        for statement in node.body:
            if isinstance(statement, ast.Expr) and isinstance(statement.value, ast.Constant) \
                    and isinstance(statement.value.value, str):
                continue                              # docstring
            if isinstance(statement, ast.Pass):
                continue
            body.append(self.visit(statement))
        return KernelFunction(params=params, body=body)

    def visit_For(self, node):
        body = [self.visit(s) for s in node.body]
        call = node.iter
        if isinstance(call, ast.Call) and isinstance(call.func, ast.Attribute) \
                and isinstance(node.target, ast.Name):
            if call.func.attr == 'interior_points':
                return InteriorPointsLoop(target=node.target.id, body=body)
            if call.func.attr == 'neighbors':
                neighbor_id = 0
                if len(call.args) >= 2:
                    id_node = call.args[1]
                    if not (isinstance(id_node, ast.Constant) and isinstance(id_node.value, int)):
                        raise StencilException(
                            "neighbors(): the neighborhood id must be an integer literal")
                    neighbor_id = id_node.value
                return NeighborPointsLoop(neighbor_id=neighbor_id,
                                          neighbor_target=node.target.id, body=body)
        raise StencilException(
            "line {}: only 'for p in self.interior_points(g)' and "
            "'for n in self.neighbors(p, k)' loops can be specialized".format(node.lineno))

    # ---- statements -----------------------------------------------------------------------
    def visit_Assign(self, node):
        if len(node.targets) != 1:
            raise StencilException("line {}: chained assignment is not supported".format(node.lineno))
        return Assign(target=self.visit(node.targets[0]), value=self.visit(node.value))

    def visit_AugAssign(self, node):
        op = _BIN_OPS.get(type(node.op))
        if op is None:
            raise StencilException("line {}: unsupported augmented operator".format(node.lineno))
        return AugAssign(target=self.visit(node.target), op=op, value=self.visit(node.value))

    def visit_Expr(self, node):
        raise StencilException("line {}: bare expressions have no effect in a stencil kernel"
                               .format(node.lineno))

    # ---- expressions ----------------------------------------------------------------------
    def visit_Constant(self, node):
        if isinstance(node.value, bool) or not isinstance(node.value, (int, float)):
            raise StencilException("only int and float literals are supported, got {!r}"
                                   .format(node.value))
        return Constant(node.value)

    def visit_Name(self, node):
        return SymbolRef(node.id)

    def visit_Attribute(self, node):
        if isinstance(node.value, ast.Name) and node.value.id == 'self':
            return SelfAttribute(node.attr)
        raise StencilException("unsupported attribute access '{}'".format(ast.dump(node)))

    def visit_BinOp(self, node):
        op = _BIN_OPS.get(type(node.op))
        if op is None:
            raise StencilException("line {}: unsupported operator {}".format(
                node.lineno, type(node.op).__name__))
        return BinaryOp(op=op, left=self.visit(node.left), right=self.visit(node.right))

    def visit_UnaryOp(self, node):
        op = _UNARY_OPS.get(type(node.op))
        if op is None:
            raise StencilException("unsupported unary operator {}".format(type(node.op).__name__))
        operand = self.visit(node.operand)
        if isinstance(operand, Constant):             # -4 is a literal, as it prints in C
            return Constant(-operand.value if op == '-' else operand.value)
        return UnaryOp(op=op, operand=operand)

    def visit_Call(self, node):
        if node.keywords:
            raise StencilException("keyword arguments are not supported in kernel calls")
        args = [self.visit(a) for a in node.args]
        if isinstance(node.func, ast.Name):
            func_name = node.func.id
        elif isinstance(node.func, ast.Attribute):
            func_name = node.func.attr
        else:
            raise StencilException("unsupported call target")
        if func_name in ('distance', 'int'):          # reference :71-73
            return MathFunction(func=func_name, args=args)
        return FunctionCall(func=func_name, args=args)

    def visit_Subscript(self, node):
        if not isinstance(node.value, ast.Name):
            raise StencilException("only plain grid names can be subscripted")
        return GridElement(grid_name=node.value.id, target=self.visit(node.slice))

    def generic_visit(self, node):
        raise StencilException("unsupported Python construct in stencil kernel: {}".format(
            type(node).__name__))
