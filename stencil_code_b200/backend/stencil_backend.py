"""StencilBackend -- the backend-independent half of specialisation.

Mirror of reference ``stencil_code/backend/stencil_backend.py:11-130``:

  * the constructor takes ``(input_grids, output_grid, kernel, arg_cfg=, fusable_nodes=,
    testing=)`` and refuses subclasses without ``visit_InteriorPointsLoop`` (``:12-17``);
  * it consumes ``kernel.ghost_depth / is_clamped / is_copied / constants / distance / dim /
    neighbors`` (``:23-35,85``);
  * parameters map to grids by position, the last parameter is the output (``:41-53``);
  * every ``NeighborPointsLoop`` is unrolled: one copy of the body per offset, in neighborhood
    order, offsets obtained from ``kernel.neighbors(origin, id)`` (``:67-90``);
  * ``self.distance(..)`` folds to ``Constant(kernel.distance(origin, offset))`` evaluated in
    Python, ``int(e)`` becomes a C cast (``:116-121``);
  * names found in ``kernel.constants`` become literals (``:127-130``).

Where the reference emits ctree C nodes, this class emits the typed ``PointProgram``
(``point_program.py``); concrete back ends turn that into code in ``visit_InteriorPointsLoop``.
"""
import numbers

import numpy

from ..stencil_exception import StencilException
from ..stencil_model import InteriorPointsLoop, GridElement, SymbolRef
from . import point_program as pp


class StencilBackend(object):
    def __init__(self, input_grids=None, output_grid=None, kernel=None, arg_cfg=None,
                 fusable_nodes=None, testing=False):
        if not hasattr(self, "visit_InteriorPointsLoop"):
            raise StencilException(
                "Error: {} must define a visit_InteriorPointsLoop method".format(type(self)))
        self.input_grids = input_grids
        self.output_grid = output_grid
        self.kernel = kernel
        self.ghost_depth = kernel.ghost_depth
        self.is_clamped = kernel.is_clamped
        self.is_copied = kernel.is_copied
        self.constants = kernel.constants
        self.distance = kernel.distance
        self.arg_cfg = arg_cfg
        self.fusable_nodes = fusable_nodes
        self.testing = testing

        self.input_dict = {}          # parameter name -> ArgInfo
        self.input_names = []
        self.output_grid_name = None
        self.kernel_target = None
        self.neighbor_target = None
        self.offset_list = None
        self.locals = {}              # temp name -> ctype

    # ---- dispatch ---------------------------------------------------------------------------
    def visit(self, node):
        method = getattr(self, 'visit_' + type(node).__name__, None)
        if method is None:
            raise StencilException("{} cannot lower a {} node".format(
                type(self).__name__, type(node).__name__))
        return method(node)

    # ---- function level (reference visit_FunctionDecl :41-53) -------------------------------
    def visit_KernelFunction(self, node):
        n_inputs = len(self.arg_cfg)
        if len(node.params) != n_inputs + 1:
            raise StencilException(
                "kernel takes {} grids (inputs + output) but was called with {} inputs".format(
                    len(node.params), n_inputs))
        for index, name in enumerate(node.params[:-1]):
            cfg = self.arg_cfg[index]
            self.input_dict[name] = pp.ArgInfo(name, index, cfg.shape, cfg.dtype)
            self.input_names.append(name)
        self.output_grid_name = node.params[-1]
        loops = [s for s in node.body if isinstance(s, InteriorPointsLoop)]
        if len(loops) != 1 or len(node.body) != 1:
            raise StencilException(
                "a stencil kernel must consist of exactly one 'for p in self.interior_points(g)' "
                "loop")
        return self.visit(loops[0])

    def lower_interior_body(self, node):
        """Unrolled, typed statement list of an InteriorPointsLoop body."""
        self.kernel_target = node.target
        statements = []
        for statement in node.body:
            lowered = self.visit(statement)
            statements.extend(lowered if isinstance(lowered, list) else [lowered])
        self.kernel_target = None
        return statements

    @property
    def out_ctype(self):
        return pp.ctype_of_dtype(self.arg_cfg[0].dtype)

    # ---- neighbor unrolling (reference :67-90) ----------------------------------------------
    def visit_NeighborPointsLoop(self, node):
        if self.kernel_target is None:
            raise StencilException("neighbors() loop outside of an interior_points() loop")
        if self.neighbor_target is not None:
            raise StencilException("nested neighbors() loops cannot be specialized")
        zero_point = tuple(0 for _ in range(self.kernel.dim))
        self.neighbor_target = node.neighbor_target
        saved_shape = getattr(self.kernel, 'current_shape', None)
        self.kernel.current_shape = None          # offsets must come out un-clamped (:85)
        body = []
        try:
            for absolute in self.kernel.neighbors(zero_point, node.neighbor_id):
                self.offset_list = [int(c) for c in absolute]
                for statement in node.body:
                    lowered = self.visit(statement)
                    body.extend(lowered if isinstance(lowered, list) else [lowered])
        finally:
            self.kernel.current_shape = saved_shape
            self.neighbor_target = None
            self.offset_list = None
        return body

    # ---- statements -------------------------------------------------------------------------
    def _store(self, target, op, value):
        if isinstance(target, GridElement):
            if target.grid_name != self.output_grid_name:
                raise StencilException("only the output grid '{}' may be assigned, not '{}'"
                                       .format(self.output_grid_name, target.grid_name))
            if not (isinstance(target.target, SymbolRef)
                    and target.target.name == self.kernel_target):
                raise StencilException("the output grid must be indexed by the interior point "
                                       "'{}'".format(self.kernel_target))
            return pp.Store(op, self.visit(value))
        if isinstance(target, SymbolRef):
            expr = self.visit(value)
            name = target.name
            if name in self.input_dict or name in (self.output_grid_name, self.kernel_target):
                raise StencilException("cannot assign to '{}'".format(name))
            declares = name not in self.locals
            if declares:
                if op != '=':
                    raise StencilException("'{}' used before assignment".format(name))
                self.locals[name] = expr.ctype
            ctype = self.locals[name]
            return pp.LocalAssign('_t_' + name, op, expr, ctype, declares)
        raise StencilException("unsupported assignment target")

    def visit_Assign(self, node):
        return self._store(node.target, '=', node.value)

    def visit_AugAssign(self, node):
        return self._store(node.target, node.op + '=', node.value)

    # ---- expressions ------------------------------------------------------------------------
    def _literal(self, value, what):
        if isinstance(value, (bool, numpy.bool_)):
            return pp.Const(int(value))
        if isinstance(value, (numbers.Integral, numpy.integer)):
            return pp.Const(int(value))
        if isinstance(value, (numbers.Real, numpy.floating)):
            return pp.Const(float(value))
        raise StencilException("{} evaluates to {!r}, which is neither int nor float"
                               .format(what, value))

    def visit_Constant(self, node):
        return self._literal(node.value, "literal")

    def visit_SymbolRef(self, node):              # reference :127-130
        if node.name in self.constants:
            return self._literal(self.constants[node.name], "constant '{}'".format(node.name))
        if node.name in self.locals:
            return pp.LocalRef('_t_' + node.name, self.locals[node.name])
        raise StencilException(
            "name '{}' is not a grid index, a local or an entry of the stencil's 'constants'"
            .format(node.name))

    def visit_SelfAttribute(self, node):
        # additive: self.<name> folds like a constants entry (the reference has no equivalent)
        if node.name in self.constants:
            return self._literal(self.constants[node.name], "constant '{}'".format(node.name))
        try:
            value = getattr(self.kernel, node.name)
        except AttributeError:
            raise StencilException("stencil has no attribute '{}'".format(node.name))
        return self._literal(value, "self.{}".format(node.name))

    def visit_BinaryOp(self, node):
        return pp.Bin(node.op, self.visit(node.left), self.visit(node.right))

    def visit_UnaryOp(self, node):
        operand = self.visit(node.operand)
        return pp.Neg(operand) if node.op == '-' else operand

    def visit_FunctionCall(self, node):
        return pp.Call(node.func, [self.visit(a) for a in node.args])

    def visit_MathFunction(self, node):           # reference :116-121
        if str(node.func) == 'distance':
            if self.offset_list is None:
                raise StencilException("self.distance() is only meaningful inside a neighbors() loop")
            zero_point = tuple(0 for _ in self.offset_list)
            return self._literal(self.distance(zero_point, tuple(self.offset_list)),
                                 "self.distance()")
        if str(node.func) == 'int':
            if len(node.args) != 1:
                raise StencilException("int() takes one argument")
            return pp.Cast(pp.INT, self.visit(node.args[0]))
        raise StencilException("unknown math function {}".format(node.func))

    def visit_GridElement(self, node):
        """Grid and table references (reference ``backend/c.py:139-188``)."""
        grid_name = node.grid_name
        target = node.target
        if isinstance(target, SymbolRef) and target.name not in self.locals \
                and target.name not in self.constants:
            name = target.name
            if name == self.kernel_target:
                if grid_name == self.output_grid_name:
                    return pp.OutRef(self.out_ctype)
                if grid_name in self.input_dict:
                    arg = self.input_dict[grid_name]
                    self._check_grid(arg)
                    return pp.Tap(arg.index, (0,) * self.kernel.dim, arg.ctype)
            else:
                if name != self.neighbor_target or self.offset_list is None:
                    raise StencilException(
                        "grid index '{}' is neither the interior point nor the neighbor of an "
                        "enclosing neighbors() loop".format(name))
                if grid_name in self.input_dict:
                    arg = self.input_dict[grid_name]
                    self._check_grid(arg)
                    return pp.Tap(arg.index, self.offset_list, arg.ctype)
                if grid_name == self.output_grid_name:
                    raise StencilException("reading neighbors of the output grid is a race and "
                                           "cannot be specialized")
            raise StencilException("unknown grid '{}'".format(grid_name))
        # computed index: a 1-d table lookup (reference :184-186)
        if grid_name not in self.input_dict:
            raise StencilException("unknown table '{}'".format(grid_name))
        arg = self.input_dict[grid_name]
        index = self.visit(target)
        if index.ctype not in (pp.INT, pp.LONG):
            raise StencilException("table '{}' must be indexed by an integer expression "
                                   "(wrap it in int())".format(grid_name))
        arg.used_as_table = True
        return pp.TableRef(arg.index, index, arg.ctype)

    def _check_grid(self, arg):
        if len(arg.shape) != self.kernel.dim:
            raise StencilException(
                "grid '{}' has {} dimensions but the stencil's neighborhoods have {}".format(
                    arg.name, len(arg.shape), self.kernel.dim))
        if arg.shape != tuple(self.arg_cfg[0].shape):
            raise StencilException(
                "grid '{}' has shape {} but the first argument has shape {}; all grids of one "
                "call must share a shape".format(arg.name, arg.shape, tuple(self.arg_cfg[0].shape)))
        arg.used_as_grid = True
