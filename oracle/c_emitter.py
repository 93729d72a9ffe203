"""Restatement of the reference's C / OpenMP code generators as a text emitter.

Follows, statement by statement:
  * ``backend/c.py:9-40``   prologue: ``abs`` prototype, ``min`` and ``clamp`` macros, one
                            ``_<grid>_array_macro`` per argument from its strides, clock timing
  * ``backend/c.py:42-137`` loop nest (clamp/copy: every point; zero/wrap: interior only),
                            ``int xN = _out_array_macro(...)``, the copy-mode halo ``if``
  * ``backend/c.py:139-188`` grid references, index-wise ``clamp`` with the grid's own shape
  * ``backend/stencil_backend.py:67-90``  neighbor loop unrolling in neighborhood order
  * ``backend/stencil_backend.py:116-130`` ``distance`` -> literal, ``int`` -> cast, constants
  * ``backend/omp.py:42-107`` the OpenMP variant: interior-only loops for every boundary mode,
                            ``#pragma omp parallel for`` in front of loop ``dim-2`` only when that
                            loop is not the outermost one, ``#pragma ivdep`` on the innermost
  * ``python_frontend.py:42-84`` what the front end recognises

It walks the Python AST of ``kernel`` itself and shares no code with the product's front end or
back end, so a bug there cannot hide here.  Works on any object that offers the reference's
``Stencil`` attributes (``kernel``, ``neighbors``, ``distance``, ``constants``, ``ghost_depth``,
``is_clamped``, ``is_copied``, ``dim``) -- the product's class and the reference's alike.
"""
import ast
import inspect
import numbers
import textwrap

import numpy

C_TYPES = {numpy.dtype('float32'): 'float', numpy.dtype('float64'): 'double',
           numpy.dtype('int32'): 'int', numpy.dtype('int64'): 'long long'}

VARIANTS = ('c', 'omp', 'c_parallel')


class OracleError(Exception):
    pass


def _literal(value):
    """How ctree's Constant node prints a Python / numpy number (assumed ``str(value)``)."""
    if isinstance(value, (bool, numpy.bool_)):
        return str(int(value))
    if isinstance(value, (numbers.Integral, numpy.integer)):
        return str(int(value))
    if isinstance(value, (numbers.Real, numpy.floating)):
        return repr(float(value))
    raise OracleError("cannot print {!r} as a C literal".format(value))


class _BodyEmitter(ast.NodeVisitor):
    """Prints the statements of the interior loop body for one specialisation."""

    def __init__(self, stencil, params, shapes, variant):
        self.stencil = stencil
        self.inputs = params[:-1]
        self.output = params[-1]
        self.shapes = dict(zip(params, shapes))
        self.variant = variant
        # the OpenMP generator never clamps (backend/omp.py:84-102)
        self.clamped = bool(stencil.is_clamped) and variant != 'omp'
        self.loop_vars = []
        self.out_index = None
        self.point = None          # interior loop target
        self.neighbor = None       # neighbor loop target
        self.offset = None

    # -- expressions --------------------------------------------------------------------------
    def visit_Constant(self, node):
        return _literal(node.value)

    def visit_Name(self, node):
        constants = self.stencil.constants
        if node.id in constants:                       # stencil_backend.py:127-130
            return _literal(constants[node.id])
        return node.id

    def visit_Attribute(self, node):
        if isinstance(node.value, ast.Name) and node.value.id == 'self':
            constants = self.stencil.constants
            if node.attr in constants:
                return _literal(constants[node.attr])
            return _literal(getattr(self.stencil, node.attr))
        raise OracleError("unsupported attribute")

    def visit_BinOp(self, node):
        op = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}[type(node.op)]
        return "({} {} {})".format(self.visit(node.left), op, self.visit(node.right))

    def visit_UnaryOp(self, node):
        op = {ast.USub: '-', ast.UAdd: '+'}[type(node.op)]
        return "({}{})".format(op, self.visit(node.operand))

    def visit_Call(self, node):
        name = node.func.id if isinstance(node.func, ast.Name) else node.func.attr
        if name == 'distance':                         # stencil_backend.py:116-119
            zero = tuple(0 for _ in self.offset)
            return _literal(self.stencil.distance(zero, tuple(self.offset)))
        if name == 'int':                              # stencil_backend.py:120-121
            return "((int){})".format(self.visit(node.args[0]))
        return "{}({})".format(name, ", ".join(self.visit(a) for a in node.args))

    def _clamp(self, text, upper):                     # c.py:147-148
        return "clamp({}, 0, {})".format(text, upper)

    def visit_Subscript(self, node):
        grid = node.value.id
        index = node.slice
        if isinstance(index, ast.Name) and index.id not in self.stencil.constants:
            if index.id == self.point:
                if grid == self.output:
                    return "{}[{}]".format(grid, self.out_index)
                shape = self.shapes[grid]
                if self.clamped:                       # c.py:160-166
                    point = [self._clamp(v, shape[d] - 1) for d, v in enumerate(self.loop_vars)]
                else:
                    point = list(self.loop_vars)
            else:                                      # c.py:170-183
                shape = self.shapes[grid]
                shifted = ["{} + {}".format(v, o) for v, o in zip(self.loop_vars, self.offset)]
                if self.clamped:
                    point = [self._clamp(s, shape[d] - 1) for d, s in enumerate(shifted)]
                else:
                    point = shifted
            return "{g}[_{g}_array_macro({p})]".format(g=grid, p=", ".join(point))
        return "{}[{}]".format(grid, self.visit(index))      # c.py:184-186

    # -- statements ---------------------------------------------------------------------------
    def visit_Assign(self, node):
        return ["{} = {};".format(self.visit(node.targets[0]), self.visit(node.value))]

    def visit_AugAssign(self, node):
        op = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}[type(node.op)]
        return ["{} {}= {};".format(self.visit(node.target), op, self.visit(node.value))]

    def visit_For(self, node):
        call = node.iter
        attr = call.func.attr
        if attr == 'neighbors':                        # stencil_backend.py:67-90
            neighbor_id = call.args[1].value if len(call.args) > 1 else 0
            zero = tuple(0 for _ in range(self.stencil.dim))
            saved = getattr(self.stencil, 'current_shape', None)
            self.stencil.current_shape = None
            self.neighbor = node.target.id
            lines = []
            try:
                for absolute in self.stencil.neighbors(zero, neighbor_id):
                    self.offset = [int(c) for c in absolute]
                    for statement in node.body:
                        lines.extend(self.visit(statement))
            finally:
                self.stencil.current_shape = saved
                self.neighbor = self.offset = None
            return lines
        raise OracleError("unexpected loop over self.{}".format(attr))

    def generic_visit(self, node):
        raise OracleError("oracle cannot restate {}".format(type(node).__name__))


def emit_c(stencil, arg_shapes, arg_dtypes, variant='c'):
    """C source of ``stencil_kernel`` for inputs of the given shapes/dtypes.

    variant 'c'          -- reference C backend (serial; the parity target)
            'omp'        -- reference OpenMP backend as written (interior only, its pragma placement)
            'c_parallel' -- C-backend semantics plus ``#pragma omp parallel for`` on the outermost
                            loop (not a reference backend; a fair all-cores baseline for 2-d grids)
    """
    if variant not in VARIANTS:
        raise OracleError("unknown variant " + variant)
    source = textwrap.dedent(inspect.getsource(stencil.kernel))
    function = [n for n in ast.parse(source).body if isinstance(n, ast.FunctionDef)][0]
    params = [a.arg for a in function.args.args][1:]            # python_frontend.py:20-22
    if len(params) != len(arg_shapes) + 1:
        raise OracleError("kernel expects {} inputs, got {}".format(len(params) - 1, len(arg_shapes)))
    out_shape = tuple(arg_shapes[0])
    out_dtype = numpy.dtype(arg_dtypes[0])
    shapes = [tuple(s) for s in arg_shapes] + [out_shape]
    dtypes = [numpy.dtype(d) for d in arg_dtypes] + [out_dtype]
    dim = len(out_shape)
    ghost = stencil.ghost_depth

    loops = [n for n in function.body if isinstance(n, ast.For)]
    if len(loops) != 1 or loops[0].iter.func.attr != 'interior_points':
        raise OracleError("kernel must be one interior_points loop")
    interior = loops[0]

    emitter = _BodyEmitter(stencil, params, shapes, variant)
    emitter.point = interior.target.id
    all_points = (stencil.is_clamped or stencil.is_copied) and variant != 'omp'   # c.py:61-81
    lines = []
    indent = "    "
    for d in range(dim):
        var = "x{}".format(d + 1)
        emitter.loop_vars.append(var)
        lo, hi = (0, out_shape[d] - 1) if all_points else (ghost[d], out_shape[d] - ghost[d] - 1)
        if variant == 'omp' and d >= 1:                 # omp.py:61-68
            if d == dim - 2:
                lines.append(indent * (d + 1) + "#pragma omp parallel for")
            elif d == dim - 1:
                lines.append(indent * (d + 1) + "#pragma ivdep")
        if variant == 'c_parallel' and d == 0:
            lines.append(indent + "#pragma omp parallel for")
        lines.append("{}for (int {v} = {lo}; {v} <= {hi}; {v}++) {{".format(
            indent * (d + 1), v=var, lo=lo, hi=hi))
    body_indent = indent * (dim + 1)
    emitter.out_index = "x{}".format(dim + 1)
    lines.append("{}int {} = _{}_array_macro({});".format(
        body_indent, emitter.out_index, params[-1], ", ".join(emitter.loop_vars)))

    statements = []
    for statement in interior.body:
        statements.extend(emitter.visit(statement))

    if stencil.is_copied and variant != 'omp':          # c.py:94-128
        tests = ["({v} < {g} || {v} > {h})".format(v=v, g=ghost[d], h=out_shape[d] - (ghost[d] + 1))
                 for d, v in enumerate(emitter.loop_vars)]
        lines.append("{}if ({}) {{".format(body_indent, " || ".join(tests)))
        lines.append("{}    {out}[{i}] = {inp}[{i}];".format(
            body_indent, out=params[-1], inp=params[0], i=emitter.out_index))
        lines.append(body_indent + "} else {")
        lines.extend(body_indent + indent + s for s in statements)
        lines.append(body_indent + "}")
    else:
        lines.extend(body_indent + s for s in statements)
    for d in reversed(range(dim)):
        lines.append(indent * (d + 1) + "}")

    head = []
    parallel = variant in ('omp', 'c_parallel')
    head.append("#include <omp.h>" if parallel else "#include <time.h>")
    head.append("int abs(int n);")                                              # c.py:22-24
    head.append("#define min(_a, _b) (((_a) < (_b)) ? (_a) : (_b))")            # c.py:25-27
    head.append("#define clamp(_a,_min_a,_max_a) "
                "(_a>_max_a?_max_a:_a)<_min_a?_min_a:(_a>_max_a?_max_a:_a)")    # c.py:28-31
    for name, shape in zip(params, shapes):                                     # c.py:11-21
        ndim = len(shape)
        strides = [int(numpy.prod(shape[k + 1:], dtype=numpy.int64)) for k in range(ndim)]
        calc = "((_d{})".format(ndim - 1)
        for k in range(ndim - 1):
            calc += "+((_d{}) * {})".format(k, strides[k])
        calc += ")"
        head.append("#define _{}_array_macro({}) {}".format(
            name, ",".join("_d{}".format(k) for k in range(ndim)), calc))
    signature = ", ".join("{}* {}".format(C_TYPES[dt], name) for name, dt in zip(params, dtypes))
    head.append("int stencil_kernel({}, float* duration) {{".format(signature))
    if parallel:                                                                # omp.py:34-39
        head.append(indent + "double start_time = omp_get_wtime();")
        tail = [indent + "*duration = omp_get_wtime() - start_time;"]
    else:                                                                       # c.py:32-39
        head.append(indent + "clock_t start_time = clock();")
        tail = [indent + "*duration = (clock() - start_time) / CLOCKS_PER_SEC;"]
    tail += [indent + "return 0;", "}"]
    return "\n".join(head + lines + tail) + "\n"
