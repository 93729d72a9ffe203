"""Strategy 'direct': one thread per grid point, taps read straight from global memory.

The always-correct generator: any dimensionality, any dtype, any shape (no alignment
requirements), all boundary modes, slab domains.  It is what small and odd-shaped grids run on
and what the tiled strategies are checked against.  Semantics are those of the reference's C
backend (``backend/c.py:42-188``): 'clamp' visits every point and clamps each tap index per
dimension; 'zero' (and 'wrap', which the reference treats identically) computes the interior and
leaves the halo 0; 'copy' computes the interior and copies the halo from the first argument.
Every output element in the requested axis-0 range is written exactly once, so the output buffer
needs no zero-fill beforehand (the reference's ``zeros`` allocation, ``stencil_kernel.py:412-416``,
is folded into the kernel: the accumulator starts at 0).
"""
import os

from ...stencil_exception import StencilException
from .codegen_common import PROLOGUE, ctype_zero, emit_statements

BLOCK_THREADS = 256


def generate(program, domain, config=None):
    if config is not None and config.get('kind') == 'tiled':
        return generate_tiled(program, domain, config)
    vector = config['V'] if config is not None and config.get('kind') == 'vector1d' else 0
    dim = program.dim
    shape = program.shape
    big = program.npoints >= 2 ** 31 - 2 ** 20
    index_t = 'i64' if big else 'int'
    strides = [1] * dim
    for d in range(dim - 2, -1, -1):
        strides[d] = strides[d + 1] * shape[d + 1]
    clamped = program.boundary == 'clamp'
    wrapped = program.boundary == 'periodic'

    def tap_text(tap):
        name = "a{}".format(tap.arg)
        if not (clamped or wrapped):
            flat = sum(o * s for o, s in zip(tap.offset, strides))
            if flat == 0:
                return "{}[idx]".format(name)
            return "{}[idx + ({}{})]".format(name, flat, 'LL' if big else '')
        terms = []
        for d in range(dim):
            o = tap.offset[d]
            if o == 0:
                coordinate = "i{}".format(d)
            elif wrapped:
                coordinate = "SC_WRAP(i{d} + ({o}), {n})".format(d=d, o=o, n=shape[d])
            else:
                coordinate = "SC_CLAMP(i{d} + ({o}), {lo}, {hi})".format(
                    d=d, o=o, lo=domain.clamp_lo[d], hi=domain.clamp_hi[d])
            if strides[d] == 1:
                terms.append("({}){}".format(index_t, coordinate))
            else:
                terms.append("({}){} * {}".format(index_t, coordinate, strides[d]))
        return "{}[{}]".format(name, " + ".join(terms))

    def table_text(ref, index_text):
        return "a{}[{}]".format(ref.arg, index_text)

    params = []
    for arg in program.args:
        params.append("const {}* __restrict__ a{}".format(arg.ctype, arg.index))
    params.append("{}* __restrict__ out".format(program.out_ctype))
    params.append("{}* __restrict__ out_peer".format(program.out_ctype))   # neighbour slab's ghost planes (or NULL)
    params.append("i64 peer_shift")                                        # element offset into out_peer
    params.append("int a0_begin")
    params.append("int a0_end")

    per0 = 1
    for s in shape[1:]:
        per0 *= s

    if domain.is_pitched:
        raise StencilException("the one-thread-per-point kernel addresses dense rows only")
    lines = [PROLOGUE]
    lines.append('extern "C" __global__ void __launch_bounds__({})'.format(BLOCK_THREADS))
    lines.append("stencil_direct({})".format(", ".join(params)))
    lines.append("{")
    lines.append("    pdl_launch_dependents();")
    lines.append("    pdl_wait();")
    lines.append("    const i64 t = (i64)blockIdx.x * {} + threadIdx.x;".format(BLOCK_THREADS))
    leave = "return"
    if vector:
        # rank 1, large grids: a thread owns V consecutive points (16 bytes).  Groups that lie V points
        # inside the grid take three 128-bit loads and one 128-bit store (no index clamping can bind,
        # no halo point is among them); the groups at either end, ranges that do not start on a
        # 16-byte boundary and the tail run the one-point code below, point by point.
        T, V, n = program.out_ctype, vector, shape[0]
        vec_t = "float4" if V == 4 else "double2"
        lines.append("    const i64 first = (i64)a0_begin + t * {};".format(V))
        lines.append("    if (first >= a0_end) return;")
        lines.append("    const int base = (int)first;")
        lines.append("    const bool aligned = ((((unsigned long long)a0) | ((unsigned long long)out)) & 15ull) == 0 "
                     "&& (a0_begin % {}) == 0;     // (a view into a larger tensor may start anywhere)".format(V))
        lines.append("    if (aligned && base >= {v} && base + {w} <= {n} && base + {v} <= a0_end) {{".format(
            v=V, w=2 * V, n=n))
        for name, where in (("qm", "base - {}".format(V)), ("q0", "base"), ("qp", "base + {}".format(V))):
            lines.append("        const {vt} {nm} = *reinterpret_cast<const {vt}*>(a0 + {w});".format(vt=vec_t, nm=name, w=where))
        values = {}
        for k, name in enumerate(("qm", "q0", "qp")):
            for c in range(V):
                values[(k - 1) * V + c] = "{}.{}".format(name, "xyzw"[c])
        for j in range(V):
            lines.append("        {} acc{} = {};".format(T, j, ctype_zero(T)))
            body = emit_statements(program, lambda tap, j=j: values[j + tap.offset[0]], table_text,
                                   acc="acc{}".format(j), indent="        ")
            lines.extend(body)
        make = "make_float4" if V == 4 else "make_double2"
        lines.append("        *reinterpret_cast<{vt}*>(out + base) = {mk}({a});".format(
            vt=vec_t, mk=make, a=", ".join("acc{}".format(j) for j in range(V))))
        lines.append("        return;")
        lines.append("    }")
        lines.append("    for (int k = 0; k < {}; ++k) {{".format(V))
        lines.append("    const int i0 = base + k;")
        lines.append("    if (i0 >= a0_end) break;")
        leave = "continue"
    else:
        lines.append("    if (t >= (i64)(a0_end - a0_begin) * {}LL) return;".format(per0))
        if dim == 1:
            lines.append("    const int i0 = a0_begin + (int)t;")
        else:
            lines.append("    const int i0 = a0_begin + (int)(t / {}LL);".format(per0))
            lines.append("    {it} rem = ({it})(t % {p}LL);".format(it='i64' if per0 >= 2 ** 31 else 'int', p=per0))
            for d in range(1, dim):
                if d < dim - 1:
                    lines.append("    const int i{d} = (int)(rem / {s}); rem = rem % {s};".format(d=d, s=strides[d]))
                else:
                    lines.append("    const int i{d} = (int)rem;".format(d=d))
    flat = " + ".join("({}){}".format(index_t, "i{}".format(d)) if strides[d] == 1
                      else "({})i{} * {}".format(index_t, d, strides[d]) for d in range(dim))
    lines.append("    const {} idx = {};".format(index_t, flat))
    if not (clamped or wrapped):
        tests = []
        for d in range(dim):
            lo, hi = domain.interior_lo[d], domain.interior_hi[d]
            if lo > 0:
                tests.append("i{} < {}".format(d, lo))
            if hi < shape[d]:
                tests.append("i{} >= {}".format(d, hi))
        if tests:
            halo_value = "a0[idx]" if program.boundary == 'copy' else ctype_zero(program.out_ctype)
            if program.boundary == 'copy' and program.args[0].ctype != program.out_ctype:
                halo_value = "({}){}".format(program.out_ctype, halo_value)
            mirror = " if (out_peer) out_peer[idx + peer_shift] = {};".format(halo_value) if domain.is_slab else ""
            lines.append("    if ({}) {{ out[idx] = {};{} {}; }}".format(" || ".join(tests), halo_value, mirror, leave))
    lines.append("    {} acc = {};".format(program.out_ctype, ctype_zero(program.out_ctype)))
    lines.extend(emit_statements(program, tap_text, table_text))
    lines.append("    out[idx] = acc;")
    if domain.is_slab:
        lines.append("    if (out_peer) out_peer[idx + peer_shift] = acc;")
    if vector:
        lines.append("    }")
    lines.append("}")
    return "\n".join(lines) + "\n"


# ---- rank 2 / rank 3: a column of points per thread ---------------------------------------------
TILE_THREADS = 128


def _tag(v):
    return "m{}".format(-v) if v < 0 else str(v)


def direct_config(program, domain):
    """How ``stencil_direct`` walks the grid.  'flat': one thread per point over the flattened
    range (any rank).  'tiled' (rank 2 and 3, rows of at least 96 elements): a thread owns RZ x RY
    points that share their last-axis index; a tap value needed by several of them is loaded once.
    The flat kernel is bound by L2 -> L1 traffic (every row is fetched once per tap row that
    touches it: 5 times for the 7-point star); the tiled one fetches 20 rows for 8 points."""
    dim = program.dim
    if dim == 1 and os.environ.get("STENCIL_CUDA_DIRECT", "tiled") != "flat":
        grids = [a for a in program.args if a.used_as_grid]
        vec = 2 if program.out_ctype == 'double' else 4
        if (program.out_ctype in ('float', 'double') and len(grids) == 1 and grids[0].index == 0
                and grids[0].ctype == program.out_ctype and program.shape[0] >= 4096
                and program.npoints < 2 ** 31 - 2 ** 20 and max(program.ghost_depth) <= vec
                and program.boundary != 'periodic' and not domain.is_slab):
            return dict(kind='vector1d', V=vec)
    if dim not in (2, 3) or program.shape[-1] < 96 or os.environ.get("STENCIL_CUDA_DIRECT", "tiled") == "flat":
        return dict(kind='flat')
    reach = max(program.ghost_depth)
    taps = {(t.arg,) + tuple(t.offset) for t in program.taps()}
    wide = 2 if program.out_ctype == 'double' else 1
    # measured on B200 (GStencil/s): 7-point (8,2) 561, (4,4) 560, (4,2) 536, (2,4) 516, (8,1) 460;
    # 27-point (4,2) 235, (8,1) 217, (4,1) 205; 5x5 (8,1) 305, (4,1) 258, (2,1) 197
    candidates = ((8, 2), (4, 2), (8, 1), (4, 1), (2, 1))
    forced = os.environ.get("STENCIL_CUDA_DIRECT_TILE")              # tuning knob: "RY,RZ"
    if forced:
        candidates = (tuple(int(v) for v in forced.split(",")),)
        wide = 0
    for RY, RZ in candidates:
        if dim == 2:
            RZ = 1
        unique = {(arg, rz + (o[0] if dim == 3 else 0), ry + o[-2], o[-1])
                  for (arg, *o) in taps for rz in range(RZ) for ry in range(RY)}
        rows = program.shape[-2]
        planes = program.shape[0] if dim == 3 else 1
        if program.boundary == 'periodic' and (rows < 2 * (RY + reach) or (dim == 3 and planes < 2 * (RZ + reach))):
            continue                      # SC_WRAP takes indices in [-n, 2n) only
        if len(unique) * wide <= 72 and -(-rows // RY) <= 65535 and -(-planes // RZ) <= 65535:
            return dict(kind='tiled', RY=RY, RZ=RZ)
    return dict(kind='flat')


def generate_tiled(program, domain, config):
    dim, shape = program.dim, program.shape
    RY, RZ = config['RY'], config['RZ']
    big = program.npoints >= 2 ** 31 - 2 ** 20
    index_t = 'i64' if big else 'int'
    n_last, n_rows = shape[-1], shape[-2]
    s_row = n_last
    s_plane = n_last * n_rows
    mode = program.boundary
    row_axis, col_axis = dim - 2, dim - 1

    def coordinate(axis, text):
        if mode == 'periodic':
            return "SC_WRAP({}, {})".format(text, shape[axis])
        if mode == 'clamp':
            return "SC_CLAMP({}, {}, {})".format(text, domain.clamp_lo[axis], domain.clamp_hi[axis])
        return "SC_CLAMP({}, 0, {})".format(text, shape[axis] - 1)      # address safety only

    taps = sorted({(t.arg,) + tuple(t.offset) for t in program.taps()})
    if mode == 'copy':
        taps = sorted(set(taps) | {(0,) + (0,) * dim})
    unique = sorted({(arg, rz + (o[0] if dim == 3 else 0), ry + o[-2], o[-1])
                     for (arg, *o) in taps for rz in range(RZ) for ry in range(RY)})
    ctype_of = {arg.index: arg.ctype for arg in program.args}

    params = []
    for arg in program.args:
        params.append("const {}* __restrict__ a{}".format(arg.ctype, arg.index))
    params += ["{}* __restrict__ out".format(program.out_ctype), "{}* __restrict__ out_peer".format(program.out_ctype),
               "i64 peer_shift", "int a0_begin", "int a0_end"]
    lines = [PROLOGUE]
    lines.append("// direct strategy, tiled: {} x {} points per thread | grid {} boundary {}".format(RZ, RY, shape, mode))
    lines.append('extern "C" __global__ void __launch_bounds__({})'.format(TILE_THREADS))
    lines.append("stencil_direct({})".format(", ".join(params)))
    lines.append("{")
    lines.append("    pdl_launch_dependents();")
    lines.append("    pdl_wait();")
    lines.append("    const int ix = blockIdx.x * {} + threadIdx.x;".format(TILE_THREADS))
    lines.append("    if (ix >= {}) return;".format(n_last))
    if dim == 3:
        lines.append("    const int iy = blockIdx.y * {};".format(RY))
        lines.append("    const int iz = a0_begin + blockIdx.z * {};".format(RZ))
        lines.append("    const int row_end = {}, plane_end = a0_end;".format(n_rows))
    else:
        lines.append("    const int iy = a0_begin + blockIdx.y * {};".format(RY))
        lines.append("    const int row_end = a0_end;")
    for c in sorted({u[3] for u in unique}):
        lines.append("    const int x_{} = {};".format(_tag(c), "ix" if c == 0 else coordinate(col_axis, "ix + ({})".format(c))))
    for c in sorted({u[2] for u in unique}):
        lines.append("    const {t} y_{n} = ({t}){e} * {s};".format(
            t=index_t, n=_tag(c), e=coordinate(row_axis, "iy + ({})".format(c)), s=s_row))
    if dim == 3:
        for c in sorted({u[1] for u in unique}):
            lines.append("    const {t} z_{n} = ({t}){e} * {s}{ll};".format(
                t=index_t, n=_tag(c), e=coordinate(0, "iz + ({})".format(c)), s=s_plane, ll='LL' if big else ''))

    def value_name(arg, cz, cy, cx):
        return "v{}_{}_{}_{}".format(arg, _tag(cz), _tag(cy), _tag(cx))
    for (arg, cz, cy, cx) in unique:
        base = "y_{}".format(_tag(cy)) + (" + z_{}".format(_tag(cz)) if dim == 3 else "")
        lines.append("    const {} {} = a{}[{} + x_{}];".format(ctype_of[arg], value_name(arg, cz, cy, cx), arg, base, _tag(cx)))

    def table_text(ref, index_text):
        return "a{}[{}]".format(ref.arg, index_text)

    halo_modes = mode in ('zero', 'copy')
    for rz in range(RZ):
        for ry in range(RY):
            valid = ["iy + {} < row_end".format(ry)]
            if dim == 3:
                valid.append("iz + {} < plane_end".format(rz))
            lines.append("    if ({}) {{".format(" && ".join(valid)))
            if dim == 3:
                lines.append("        const {t} idx = ({t})(iz + {rz}) * {sp}{ll} + ({t})(iy + {ry}) * {sr} + ix;".format(
                    t=index_t, rz=rz, ry=ry, sp=s_plane, sr=s_row, ll='LL' if big else ''))
            else:
                lines.append("        const {t} idx = ({t})(iy + {ry}) * {sr} + ix;".format(t=index_t, ry=ry, sr=s_row))
            lines.append("        {} acc = {};".format(program.out_ctype, ctype_zero(program.out_ctype)))
            tests = []
            if halo_modes:
                coords = {col_axis: "ix", row_axis: "iy + {}".format(ry)}
                if dim == 3:
                    coords[0] = "iz + {}".format(rz)
                for d in range(dim):
                    lo, hi = domain.interior_lo[d], domain.interior_hi[d]
                    if lo > 0:
                        tests.append("{} < {}".format(coords[d], lo))
                    if hi < shape[d]:
                        tests.append("{} >= {}".format(coords[d], hi))

            def tap_text(tap, rz=rz, ry=ry):
                o = tuple(tap.offset)
                return value_name(tap.arg, rz + (o[0] if dim == 3 else 0), ry + o[-2], o[-1])
            body = emit_statements(program, tap_text, table_text, indent="            ")
            if tests:
                halo_value = ctype_zero(program.out_ctype)
                if mode == 'copy':
                    halo_value = value_name(0, rz, ry, 0)
                    if program.args[0].ctype != program.out_ctype:
                        halo_value = "({}){}".format(program.out_ctype, halo_value)
                lines.append("        if ({}) acc = {};".format(" || ".join(tests), halo_value))
                lines.append("        else {")
                lines.extend(body)
                lines.append("        }")
            else:
                lines.append("        {")
                lines.extend(body)
                lines.append("        }")
            lines.append("        out[idx] = acc;")
            if domain.is_slab:
                lines.append("        if (out_peer) out_peer[idx + peer_shift] = acc;")
            lines.append("    }")
    lines.append("}")
    return "\n".join(lines) + "\n"


def launch_geometry(program, a0_begin, a0_end, config=None):
    if config is not None and config.get('kind') == 'tiled':
        tiles_x = -(-program.shape[-1] // TILE_THREADS)
        if program.dim == 3:
            grid = (tiles_x, -(-program.shape[1] // config['RY']), max(1, -(-(a0_end - a0_begin) // config['RZ'])))
        else:
            grid = (tiles_x, max(1, -(-(a0_end - a0_begin) // config['RY'])), 1)
        return grid, (TILE_THREADS, 1, 1)
    per0 = 1
    for s in program.shape[1:]:
        per0 *= s
    total = (a0_end - a0_begin) * per0
    if config is not None and config.get('kind') == 'vector1d':
        total = -(-total // config['V'])                  # a thread owns V consecutive points
    blocks = max(1, (total + BLOCK_THREADS - 1) // BLOCK_THREADS)
    return (blocks, 1, 1), (BLOCK_THREADS, 1, 1)
