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
from ...stencil_exception import StencilException
from .codegen_common import PROLOGUE, ctype_zero, emit_statements

BLOCK_THREADS = 256


def generate(program, domain):
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
            lines.append("    if ({}) {{ out[idx] = {};{} return; }}".format(" || ".join(tests), halo_value, mirror))
    lines.append("    {} acc = {};".format(program.out_ctype, ctype_zero(program.out_ctype)))
    lines.extend(emit_statements(program, tap_text, table_text))
    lines.append("    out[idx] = acc;")
    if domain.is_slab:
        lines.append("    if (out_peer) out_peer[idx + peer_shift] = acc;")
    lines.append("}")
    return "\n".join(lines) + "\n"


def launch_geometry(program, a0_begin, a0_end):
    per0 = 1
    for s in program.shape[1:]:
        per0 *= s
    total = (a0_end - a0_begin) * per0
    blocks = max(1, (total + BLOCK_THREADS - 1) // BLOCK_THREADS)
    return (blocks, 1, 1), (BLOCK_THREADS, 1, 1)
