"""Strategy 'temporal' for rank-2 grids: two sweeps fused into one pass over HBM, all in registers.

The rank-2 counterpart of ``codegen_temporal`` (iterated 2-d Jacobi / heat / image filters, and
``fuse(a, b)`` of two rank-2 stencils -- the reference's stencil-fusion experiment is 2-d:
``ipython_notebooks/fusing_stencils/Fusing Stencils.rst:8-25``).  Per launch out = S2(S1(in)) with
8 B of HBM traffic per point for TWO sweeps.

Rows are streamed along axis 0 as in ``codegen_stream`` (rank 2): a producer warp feeds a ring of
TMA stages of U rows; every consumer warp owns 128-column sub-tiles, a thread a quad of columns
per sub-tile.  What is new:

  * sweep 1 runs from the rotating input register window and leaves the intermediate ("mid") row
    in a SECOND register window; the columns next to a thread's quad are fetched from the
    neighbouring lanes with warp shuffles once, when the mid row is produced, and kept with it.
    Sweep 2 runs entirely from that window: the intermediate grid never touches shared or global
    memory.  No barrier between warps is needed: sub-tiles overlap by one quad per side (a warp
    computes mid on 128 columns and outputs the inner 120), so no warp needs another's values.
  * the loop body is emitted once per phase of lcm(W1, W2, U*S) phases: window slots, ring stage
    and row within the stage are compile-time constants (immediate LDS offsets, no index
    arithmetic, which is what bounds the single-sweep rank-2 kernel).
  * 'clamp' along axis 0 is applied to registers where a row enters a window (a uniform, rarely
    taken branch): a row past the last one takes the previous row's values, the first row is
    copied into the slots of the rows before it.  In-row 'clamp' is the usual select cascade on
    edge sub-tiles; 'zero' / 'copy' select the halo value where a row is produced.

The arithmetic of both sweeps is the unchanged per-point program: bit-identical to two sweeps.

Eligibility: rank 2, float32, one grid argument, no data-indexed tables, no local temporaries,
column reach <= 4, rows a multiple of 4 floats and not padded, not a slab.
"""
import math
import os

from ...stencil_exception import StencilException
from .. import point_program as pp
from .codegen_common import PROLOGUE, emit_statements
from .codegen_stream import PTX_HELPERS, TensorMapSpec, _Emitter, _group_loads, _grid_args, table_plan, plan_chunks, \
    _const_index

FUNCTION = "stencil_temporal2_rows"
SINGLE_FUNCTION = "stencil_rows"


def eligible(program, domain):
    if program.dim != 2 or program.out_ctype not in (pp.FLOAT, pp.DOUBLE):
        return False, "rank 2 float32 / float64 only"
    grids = _grid_args(program)
    if len(grids) != 1 or grids[0].index != 0 or grids[0].ctype != program.out_ctype:
        return False, "exactly one grid argument, of the output's type"
    vec = vector_width(program)
    if domain.is_slab or domain.is_pitched or program.shape[-1] % vec:
        return False, "slab, padded or ragged rows"
    if program.shape[-1] < 256 or program.shape[0] < 16:
        return False, "grid too small"
    if program.boundary == 'periodic':
        return False, "periodic boundaries"
    if table_plan(program, False)[0]:
        return False, "data-indexed tables"
    if any(isinstance(s, pp.LocalAssign) for s in program.statements):
        return False, "local temporaries"
    lo, hi = program.tap_extent()
    if max(-lo[1], hi[1]) > vec or hi[0] - lo[0] > 8:
        return False, "tap reach"
    return True, "rank-2 stencil"


def vector_width(program):
    """elements a thread owns per row: 16 bytes -- four floats or two doubles"""
    return 2 if program.out_ctype == pp.DOUBLE else 4


def pair_eligible(first, first_domain, second, second_domain):
    for program, domain in ((first, first_domain), (second, second_domain)):
        ok, why = eligible(program, domain)
        if not ok:
            return False, why
        if len(program.args) != 1:
            return False, "stencils with table arguments are not fused with other stencils"
    if first.out_ctype != second.out_ctype:
        return False, "element types differ"
    if first.shape != second.shape or first.boundary != second.boundary:
        return False, "shape or boundary mode differ"
    return True, "pair of rank-2 stencils"


class Temporal2dConfig(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def describe(self):
        keys = ("NCW", "BY", "U", "S", "HX", "W1", "W2", "P", "OUT_COLS", "TX", "blocks_per_sm", "smem_bytes")
        return dict({k: getattr(self, k) for k in keys}, strategy="temporal2_rows")


def _extents(program, second=None):
    second = program if second is None else second
    lo1, hi1 = program.tap_extent()
    lo2, hi2 = second.tap_extent()
    hx = max(-lo1[1], hi1[1], -lo2[1], hi2[1])
    return (lo1[0], hi1[0]), (lo2[0], hi2[0]), hx


def choose_config(program, device_props, second=None, depth=2):
    """``depth`` 2: the fused two-sweep kernel; 1: the same row pipeline with a single sweep (the
    single-sweep rank-2 kernel with compile-time ring indices, ``SINGLE_FUNCTION``)."""
    (lo1, hi1), (lo2, hi2), hx = _extents(program, second)
    if depth == 1:
        lo2 = hi2 = 0
    W1, W2 = hi1 - lo1 + 1, hi2 - lo2 + 1
    V = vector_width(program)
    elem = 8 if V == 2 else 4
    PADX = V if hx else 0
    override = {}
    knob = "STENCIL_CUDA_TEMPORAL_TILE" if depth == 2 else "STENCIL_CUDA_ROWS_TILE"
    for item in filter(None, os.environ.get(knob, "").split(",")):
        key, _, value = item.partition("=")
        override[key.strip()] = int(value)
    window = W1 * W2 // math.gcd(W1, W2)
    width = (V + 2 * hx) * (elem // 4)          # 32-bit registers of a window row
    # registers: both windows, two accumulator quads, addresses
    best = None
    # measured on B200, 16384^2 (GStencil/s).  Fused, 4-point Jacobi: (4,2) 1491, (8,1) 1477, (4,1) 1373.
    # Single sweeps, 5x5 convolution, clamp / zero / copy alike: (16,1) 772-776, (8,1) 641-753, (4,1) 677-697,
    # (4,2) 648-685 (codegen_stream's kernel: 703-710); 4-point Jacobi: (8,1) 765, (16,1) 741 (codegen_stream: 772)
    if depth == 2:
        order = ((4, 2), (4, 1), (8, 1), (2, 1))
    elif W1 >= 4:
        order = ((16, 1), (8, 1), (4, 1), (2, 1))
    else:
        order = ((8, 1), (4, 1), (4, 2), (2, 1))
    for NCW, BY in order:
        NCW, BY = override.get("NCW", NCW), override.get("BY", BY)
        registers = 32 + BY * (W1 + (W2 if depth == 2 else 0)) * width + BY * 2 * V * (elem // 4)
        if registers > 168 and not override:
            continue
        # rows per stage: the window length where that keeps the phase count small
        U = override.get("U", window if 3 <= window <= 8 else 4)
        S = override.get("S", 4 if U <= 4 else 3)
        P = window * (U * S) // math.gcd(window, U * S)
        if P > 24 and not override:
            U, S = window, 3
            P = window * (U * S) // math.gcd(window, U * S)
        pitch = 32 * V + 2 * PADX                                 # elements of a TMA box row
        n_sub = NCW * BY
        sub_floats = (U * pitch * elem + 127) // 128 * (128 // elem)   # elements; every TMA box starts 128-byte aligned
        stage_floats = n_sub * sub_floats
        smem = S * stage_floats * elem + 2 * S * 8 + 128
        threads = (NCW + 1) * 32
        if smem > device_props.smem_per_block_optin:
            continue
        blocks = min(device_props.smem_per_sm // (smem + 1024), device_props.max_threads_per_sm // threads, 8,
                     max(1, device_props.regs_per_sm // (threads * max(registers, 40))))
        if program.uses_double() or (second is not None and second.uses_double()):
            blocks = min(blocks, 2)                 # double temporaries: leave room, avoid spills
        blocks = override.get("blocks", blocks)
        out_cols = 32 * V - (2 * V if hx and depth == 2 else 0)
        best = Temporal2dConfig(DEPTH=depth, V=V, ELEM=elem, NCW=NCW, BY=BY, U=U, S=S, HX=hx, PADX=PADX, PITCH=pitch, NSUB=n_sub, W1=W1, W2=W2, P=P,
                                LO1=lo1, HI1=hi1, LO2=lo2, HI2=hi2, STAGE_FLOATS=stage_floats,
                                STAGE_BYTES=n_sub * U * pitch * elem, SUB_FLOATS=sub_floats, OUT_COLS=out_cols,
                                TX=n_sub * out_cols, smem_bytes=smem, threads=threads, blocks_per_sm=max(1, blocks))
        break
    if best is None:
        raise StencilException("no rank-2 temporal tile fits")
    return best


def worthwhile(program, cfg):
    """The fused kernel halves the traffic but computes 128 columns per 120 it outputs; it pays on
    grids that are HBM bound (far larger than L2) and wide enough to fill the sub-tiles."""
    n0, n1 = program.shape
    tiles = -(-n1 // cfg.TX)
    useful = n1 / float(tiles * cfg.NSUB * 32 * cfg.V)
    # arithmetic per point: the fused kernel is issue bound for large neighbourhoods (measured on
    # B200, 16384^2, GStencil/s fused vs single sweeps: 4-point Jacobi 1491 vs 773, 3x3 sum 1435 vs
    # 763, 5x5 convolution 615 vs 723)
    operations = sum(1 for node in program.walk() if isinstance(node, pp.Bin)) + \
        sum(1 for s in program.statements if isinstance(s, pp.Store) and s.op != '=')
    return useful >= 0.80 and n0 * n1 >= (1 << 23) and operations <= 24


def _tag(v):
    return "m{}".format(-v) if isinstance(v, int) and v < 0 else str(v)


def generate(program, domain, cfg, second=None, second_domain=None, function=FUNCTION):
    second = program if second is None else second
    second_domain = domain if second_domain is None else second_domain
    N0, N1 = program.shape
    NCW, BY, U, S, HX, PADX, PITCH = cfg.NCW, cfg.BY, cfg.U, cfg.S, cfg.HX, cfg.PADX, cfg.PITCH
    W1, W2, P = cfg.W1, cfg.W2, cfg.P
    LO1, HI1, LO2, HI2 = cfg.LO1, cfg.HI1, cfg.LO2, cfg.HI2
    SF, SUBF, OUTC = cfg.STAGE_FLOATS, cfg.SUB_FLOATS, cfg.OUT_COLS
    boundary = program.boundary
    clamp, copy = boundary == 'clamp', boundary == 'copy'
    c_lo, c_hi = domain.clamp_lo, domain.clamp_hi
    i_lo, i_hi = domain.interior_lo, domain.interior_hi
    j_lo, j_hi = second_domain.interior_lo, second_domain.interior_hi
    two_masks = (tuple(i_lo), tuple(i_hi)) != (tuple(j_lo), tuple(j_hi))
    one_sweep = cfg.DEPTH == 1                     # the row enters the window, the result is stored
    V, ELEM = cfg.V, cfg.ELEM                      # elements per thread and row (16 bytes), bytes per element
    T = program.out_ctype                          # 'float' | 'double'
    ZERO = "0.0f" if T == pp.FLOAT else "0.0"
    WIDTH = 32 * V                                 # columns of a warp's sub-tile
    LANE_PAD = V if (HX and not one_sweep) else 0  # the first / last lane of a sub-tile only produce halo
    cols = list(range(-HX, V + HX))                # columns of a window row, relative to the thread's first one
    _, const_refs, _, _ = table_plan(program, False)
    _, const_refs2, _, _ = table_plan(second, False)
    const_refs = dict(const_refs)
    const_refs.update(const_refs2)               # keys are (argument, index) pairs
    pitch = domain.row_pitch

    def in_name(slot, b, col):
        return "i{}_{}_{}".format(slot, b, _tag(col))

    def mid_name(slot, b, col):
        return "m{}_{}_{}".format(slot, b, _tag(col))

    def table_text(ref, index_text):
        return "tc{}_{}".format(ref.arg, _const_index(ref.index))

    e = _Emitter()
    e.lines.append(PROLOGUE)
    e.lines.append(PTX_HELPERS.replace("@STOP@", os.environ.get("STENCIL_CUDA_STORE_OP", ".cs")))
    if T == pp.DOUBLE:
        e.lines.append("""__device__ __forceinline__ void st_global_v2d(double* p, double a, double b) {
    asm volatile("st.global%s.v2.f64 [%%0], {%%1, %%2};" :: "l"(p), "d"(a), "d"(b) : "memory");
}""" % os.environ.get("STENCIL_CUDA_STORE_OP", ".cs"))
    store = "st_global_v4" if T == pp.FLOAT else "st_global_v2d"
    e("// temporal strategy, rank 2: two fused sweeps in registers | grid {} boundary {} | {} warps x {} "
      "sub-tiles of {} columns ({} outputs each), U={} rows per stage, ring S={}, windows {}/{}, {} phases".format(
          program.shape, boundary, NCW, BY, WIDTH, OUTC, U, S, W1, W2, P))
    params = ["const {}* __restrict__ a{}".format(arg.ctype, arg.index) for arg in program.args]
    params += ["const __grid_constant__ TensorMap tm0", "{}* __restrict__ out".format(T),
               "{}* __restrict__ out_peer".format(T), "i64 peer_shift", "int a0_begin", "int a0_end", "int chunk_len"]
    e('extern "C" __global__ void __launch_bounds__({}, {})'.format(cfg.threads, cfg.blocks_per_sm))
    e("{}({})".format(function, ", ".join(params)))
    e.open("")
    e("extern __shared__ __align__(128) unsigned char smem_raw[];")
    e("{t}* const ring = reinterpret_cast<{t}*>(smem_raw);".format(t=T))
    e("u64* const bar_full = reinterpret_cast<u64*>(smem_raw + {});".format(S * SF * ELEM))
    e("u64* const bar_empty = bar_full + {};".format(S))
    e("const int tid = threadIdx.x;")
    e("const int warp = tid >> 5;")
    e("const int lane = tid & 31;")
    e("const int x0 = blockIdx.x * {};           // first output column of the CTA".format(cfg.TX))
    e("const int u_begin = a0_begin + blockIdx.z * chunk_len;")
    e("const int u_end = min(u_begin + chunk_len, a0_end);")
    e("if (u_begin >= u_end) return;")
    e("const int p_first = u_begin + ({});        // axis-0 index of streamed row 0".format(LO1 + LO2))
    e("const int n_units = (u_end - u_begin) + {};".format((W1 - 1) + (W2 - 1)))
    e("const int n_stages = (n_units + {}) / {};".format(U - 1, U))
    e.open("if (tid == 0)")
    e.open("for (int s = 0; s < {}; ++s)".format(S))
    e("mbar_init(&bar_full[s], 1);")
    e("mbar_init(&bar_empty[s], {});".format(NCW))
    e.close()
    e("fence_barrier_init();")
    e.close()
    e("__syncthreads();")
    e("pdl_wait();                         // the grid may still be written by the preceding sweep")
    e()
    # ------------------------------------------------------------------ producer
    e.open("if (warp == {})".format(NCW))
    e("if (lane == 0) tma_prefetch_desc(&tm0);")
    e.open("for (int k = 0; k < n_stages; ++k)")
    e("const int s = k % {};".format(S))
    e("if (k >= {S}) mbar_wait(&bar_empty[s], ((k / {S}) - 1) & 1);".format(S=S))
    e.open("if (lane == 0)")
    e("mbar_expect_tx(&bar_full[s], {});".format(cfg.STAGE_BYTES))
    e("const int r0 = p_first + k * {};    // rows outside the grid are zero-filled; 'clamp' is applied to registers".format(U))
    e.open("for (int w = 0; w < {}; ++w)".format(cfg.NSUB))
    e("tma_load_2d(ring + s * {sf} + w * {sub}, &tm0, &bar_full[s], x0 + w * {oc} - {pad}, r0);".format(
        sf=SF, sub=SUBF, oc=OUTC, pad=LANE_PAD + PADX))
    e.close()
    e.close()
    e.close()
    e.lines.append("#ifdef SC_PDL")
    e("mbar_wait(&bar_full[(n_stages - 1) % {S}], ((n_stages - 1) / {S}) & 1);    // all loads have landed".format(S=S))
    e("pdl_launch_dependents();")
    e.lines.append("#endif")
    e("return;")
    e.close()
    e()
    # ------------------------------------------------------------------ consumers
    for b in range(BY):
        e("const int gx{b} = x0 + (warp * {by} + {b}) * {oc} - {pad} + lane * {v};   // first column of block {b}".format(
            b=b, by=BY, oc=OUTC, pad=LANE_PAD, v=V))
        e("const {t}* const src{b} = ring + (warp * {by} + {b}) * {sub} + {px} + lane * {v};".format(
            t=T, b=b, by=BY, sub=SUBF, px=PADX, v=V))
        lane_ok = "lane >= 1 && lane <= 30 && " if LANE_PAD else ""
        e("const bool store_ok{b} = {l}gx{b} >= 0 && gx{b} < {n};".format(b=b, l=lane_ok, n=N1))
        e("{t}* dst{b} = out + (i64)u_begin * {p} + gx{b};".format(t=T, b=b, p=pitch))
        if clamp and HX:
            e("const bool x_edge{b} = gx{b} - lane * {v} <= 0 || gx{b} - lane * {v} + {r} > {n};".format(
                b=b, v=V, r=WIDTH + HX, n=N1))
    if not clamp:
        for name, lo_, hi_ in (("halo_mask", i_lo, i_hi),) + ((("halo_mask2", j_lo, j_hi),) if two_masks else ()):
            e("u32 {} = 0;                   // bit (b*V+j): the point lies in the column halo".format(name))
            for b in range(BY):
                for j in range(V):
                    tests = []
                    if lo_[1] > 0:
                        tests.append("gx{} + {} < {}".format(b, j, lo_[1]))
                    if hi_[1] < N1:
                        tests.append("gx{} + {} >= {}".format(b, j, hi_[1]))
                    if tests:
                        e("if ({}) {} |= {}u;".format(" || ".join(tests), name, 1 << (b * V + j)))
    for (arg, index), ctype in sorted(const_refs.items()):
        e("const {} tc{}_{} = a{}[{}];".format(ctype, arg, index, arg, index))
    names = [in_name(slot, b, c) for slot in range(W1) for b in range(BY) for c in cols]
    if one_sweep:
        names += [mid_name(0, b, j) for b in range(BY) for j in range(V)]          # the accumulators
    else:
        names += [mid_name(slot, b, c) for slot in range(W2) for b in range(BY) for c in cols]
    for first in range(0, len(names), 8):
        e("{} {};".format(T, ", ".join(n + " = " + ZERO for n in names[first:first + 8])))
    counter = [0]

    def column_fix(b, names_by_col):
        """'clamp' along a row: cascade the nearest in-grid column outward"""
        for c in sorted((c for c in names_by_col if c < 0), reverse=True):
            e("if (gx{} + ({}) < 0) {} = {};".format(b, c, names_by_col[c], names_by_col[c + 1]))
        for c in sorted(c for c in names_by_col if c > V - 1):
            e("if (gx{} + {} > {}) {} = {};".format(b, c, N1 - 1, names_by_col[c], names_by_col[c - 1]))

    def group_loads(needed):
        """aligned 16-byte / 8-byte / scalar shared-memory loads covering the columns ``needed``"""
        if V == 4:
            return _group_loads(needed)
        loads = []
        for pair in sorted({c // 2 for c in needed}):
            used = sorted(c for c in needed if c // 2 == pair)
            loads.append((2 * pair, 2, used) if len(used) == 2 else (used[0], 1, used))
        return loads

    e("int unit = 0;                        // newest streamed input row, relative to p_first")
    e("u32 par = 0;")
    fills_per_pass = P // (U * S)
    e.open("while (true)")
    for phase in range(P):
        stage, row = (phase // U) % S, phase % U
        s1 = phase % W1                              # input-window slot of the newest row
        s2 = phase % W2                              # mid-window slot of the mid row produced in this phase
        e("// ---- phase {} (ring stage {} row {}, input slot {}, mid slot {})".format(phase, stage, row, s1, s2))
        e("if (unit >= n_units) break;")
        e.open("")
        if row == 0:
            fills = phase // (U * S)
            parity = str(fills % 2) if fills_per_pass % 2 == 0 else ("par" if fills % 2 == 0 else "(par ^ 1u)")
            e("mbar_wait(&bar_full[{}], {});".format(stage, parity))
        offset = stage * SF + row * PITCH
        # ---------------- the newest input row enters the window
        for b in range(BY):
            for first, width, used in group_loads(set(cols)):
                if width == 1:
                    e("{} = src{}[{}];".format(in_name(s1, b, used[0]), b, offset + first))
                    continue
                counter[0] += 1
                vec = {(4, 4): "float4", (4, 2): "float2", (2, 2): "double2"}[(V, width)]
                e("const {v} q{n} = *reinterpret_cast<const {v}*>(src{b} + {o});".format(v=vec, n=counter[0], b=b, o=offset + first))
                e(" ".join("{} = q{}.{};".format(in_name(s1, b, c), counter[0], "xyzw"[c - first]) for c in used))
        if row == U - 1:
            e("__syncwarp();")
            e("if (lane == 0) mbar_arrive(&bar_empty[{}]);".format(stage))
        if clamp and HX:
            for b in range(BY):
                e.open("if (x_edge{})".format(b))
                column_fix(b, {c: in_name(s1, b, c) for c in cols})
                e.close()
        if clamp and W1 > 1:
            # axis-0 clamp of the input, on registers: a row past the last one reads as the last one,
            # the rows before the first one as the first
            e("const int ri = p_first + unit;")
            e.open("if (ri <= {} || ri > {})".format(c_lo[0], c_hi[0]))
            e('asm volatile("" ::: "memory");')
            prev = (s1 - 1) % W1
            e.open("if (ri > {})".format(c_hi[0]))
            for b in range(BY):
                e(" ".join("{} = {};".format(in_name(s1, b, c), in_name(prev, b, c)) for c in cols))
            e.close()
            e.open("else if (ri == {})".format(c_lo[0]))
            for older in range(1, W1):
                for b in range(BY):
                    e(" ".join("{} = {};".format(in_name((s1 - older) % W1, b, c), in_name(s1, b, c)) for c in cols))
            e.close()
            e.close()

        # ---------------- sweep 1: mid row rm = p_first + unit - HI1
        e.open("if (unit >= {})".format(W1 - 1))
        e("const int rm = p_first + unit - ({});".format(HI1))
        for b in range(BY):
            e(" ".join("{} = {};".format(mid_name(s2, b, j), ZERO) for j in range(V)))
        for statement in program.statements:
            for b in range(BY):
                for j in range(V):
                    def tap_text(tap, b=b, j=j):
                        dr, dc = tap.offset
                        return in_name((s1 + dr - HI1) % W1, b, j + dc)
                    single = pp.PointProgram(2, program.shape, program.out_ctype, program.args, [statement],
                                             program.ghost_depth, boundary)
                    e(emit_statements(single, tap_text, table_text, acc=mid_name(s2, b, j), indent="")[0])
        if not clamp:
            tests = []
            if i_lo[0] > 0:
                tests.append("rm < {}".format(i_lo[0]))
            if i_hi[0] < N0:
                tests.append("rm >= {}".format(i_hi[0]))
            e("const bool rm_halo = {};".format(" || ".join(tests) if tests else "false"))
            e.open("if (rm_halo || halo_mask)")
            e('asm volatile("" ::: "memory");     // keeps this rare case a branch, not selects for everyone')
            for b in range(BY):
                for j in range(V):
                    value = in_name((s1 - HI1) % W1, b, j) if copy else ZERO
                    e("if (rm_halo || (halo_mask & {}u)) {} = {};".format(1 << (b * V + j), mid_name(s2, b, j), value))
            e.close()
        if one_sweep:
            for b in range(BY):
                e("if (store_ok{b}) {st}(dst{b}, {v});".format(b=b, st=store, v=", ".join(mid_name(s2, b, j) for j in range(V))))
                e("dst{b} += {p};".format(b=b, p=pitch))
        else:
            # the columns next to the quad come from the neighbouring lanes, once per mid row
            for b in range(BY):
                for c in cols:
                    if c < 0 or c > V - 1:
                        e("{} = __shfl_{}_sync(0xffffffffu, {}, 1);".format(
                            mid_name(s2, b, c), "up" if c < 0 else "down", mid_name(s2, b, c % V)))
            if clamp and HX:
                for b in range(BY):
                    e.open("if (x_edge{})".format(b))
                    column_fix(b, {c: mid_name(s2, b, c) for c in cols})
                    e.close()
            if clamp and W2 > 1:
                e.open("if (rm <= {} || rm > {})".format(c_lo[0], c_hi[0]))
                e('asm volatile("" ::: "memory");')
                prev = (s2 - 1) % W2
                e.open("if (rm > {})".format(c_hi[0]))
                for b in range(BY):
                    e(" ".join("{} = {};".format(mid_name(s2, b, c), mid_name(prev, b, c)) for c in cols))
                e.close()
                e.open("else if (rm == {})".format(c_lo[0]))
                for older in range(1, W2):
                    for b in range(BY):
                        e(" ".join("{} = {};".format(mid_name((s2 - older) % W2, b, c), mid_name(s2, b, c)) for c in cols))
                e.close()
                e.close()

            # ---------------- sweep 2: output row zo = rm - HI2
            e.open("if (unit >= {})".format((W1 - 1) + (W2 - 1)))
            acc = [["acc{}_{}".format(b, j) for j in range(V)] for b in range(BY)]
            for b in range(BY):
                e("{} {};".format(T, ", ".join(n + " = " + ZERO for n in acc[b])))
            for statement in second.statements:
                for b in range(BY):
                    for j in range(V):
                        def tap_text(tap, b=b, j=j):
                            dr, dc = tap.offset
                            return mid_name((s2 + dr - HI2) % W2, b, j + dc)
                        single = pp.PointProgram(2, second.shape, second.out_ctype, second.args, [statement],
                                                 second.ghost_depth, boundary)
                        e(emit_statements(single, tap_text, table_text, acc=acc[b][j], indent="")[0])
            if not clamp:
                e("const int zo = rm - ({});".format(HI2))
                mask2 = "halo_mask2" if two_masks else "halo_mask"
                tests = []
                if j_lo[0] > 0:
                    tests.append("zo < {}".format(j_lo[0]))
                if j_hi[0] < N0:
                    tests.append("zo >= {}".format(j_hi[0]))
                e("const bool zo_halo = {};".format(" || ".join(tests) if tests else "false"))
                e.open("if (zo_halo || {})".format(mask2))
                e('asm volatile("" ::: "memory");     // keeps this rare case a branch, not selects for everyone')
                for b in range(BY):
                    for j in range(V):
                        value = mid_name((s2 - HI2) % W2, b, j) if copy else ZERO
                        e("if (zo_halo || ({} & {}u)) {} = {};".format(mask2, 1 << (b * V + j), acc[b][j], value))
                e.close()
            for b in range(BY):
                e("if (store_ok{b}) {st}(dst{b}, {v});".format(b=b, st=store, v=", ".join(acc[b])))
                e("dst{b} += {p};".format(b=b, p=pitch))
            e.close()   # sweep 2
        e.close()   # sweep 1
        e("++unit;")
        e.close()
    if fills_per_pass % 2 == 1:
        e("par ^= 1u;")
    e.close()   # while
    e.close()   # kernel
    return e.text()


def make_plan(program, domain, device_props, options, second=None, second_domain=None):
    from .transformer import CudaPlan
    from .planner import Launch
    if second is None:
        ok, why = eligible(program, domain)
    else:
        ok, why = pair_eligible(program, domain, second, second_domain)
    if not ok:
        raise StencilException("temporal blocking not applicable: " + why)
    cfg = choose_config(program, device_props, second)
    source = generate(program, domain, cfg, second=second, second_domain=second_domain)
    n0, n1 = program.shape
    specs = [TensorMapSpec(0, (n1, n0), (domain.row_pitch * cfg.ELEM,), (cfg.PITCH, cfg.U),
                           dtype_code=1 if cfg.ELEM == 8 else 0, tag="temporal-rows")]
    tiles = -(-n1 // cfg.TX)
    slots = device_props.sm_count * cfg.blocks_per_sm
    warmup = (cfg.W1 - 1) + (cfg.W2 - 1)

    def geometry(a0_begin, a0_end):
        n_units = a0_end - a0_begin
        chunks = plan_chunks(n_units, tiles, slots, warmup + 1, max(16 * max(warmup, 1), 64))
        chunk_len = -(-n_units // chunks)
        chunks = -(-n_units // chunk_len)
        launch = Launch((tiles, 1, chunks), (cfg.threads, 1, 1), cfg.smem_bytes)
        launch.extra = (chunk_len,)
        return launch

    plan = CudaPlan(program, domain, 'temporal', source, FUNCTION, options, geometry,
                    tensor_maps=specs, notes={'config': cfg.describe()})
    plan.function_no_peer = None
    plan.function_step = None
    return plan


# ---- the same row pipeline with ONE sweep: the single-sweep kernel of rank-2 grids -------------------
def single_applicable(program, domain):
    """Single sweeps of rank-2 stencils take this generator (``stencil_rows``) where it applies --
    compile-time ring indices and register-side axis-0 clamp instead of the per-row index arithmetic
    that bounds ``codegen_stream``'s rank-2 kernel -- and ``codegen_stream`` otherwise (several grid
    arguments, data-indexed tables, slabs, padded rows, wide reach, large windows)."""
    if os.environ.get("STENCIL_CUDA_ROWS", "1") in ("0", ""):
        return False
    ok, _ = eligible(program, domain)
    if not ok:
        return False
    lo, hi = program.tap_extent()
    vec = vector_width(program)
    width = (vec + 2 * max(-lo[1], hi[1])) * (4 // vec)          # 32-bit registers of a window row
    return 32 + (hi[0] - lo[0] + 1) * width + 8 <= 168


def make_single_plan(program, domain, device_props, options):
    from .transformer import CudaPlan
    from .planner import Launch
    from .codegen_stream import plan_chunks_adaptive
    cfg = choose_config(program, device_props, depth=1)
    source = generate(program, domain, cfg, function=SINGLE_FUNCTION)
    n0, n1 = program.shape
    specs = [TensorMapSpec(0, (n1, n0), (domain.row_pitch * cfg.ELEM,), (cfg.PITCH, cfg.U),
                           dtype_code=1 if cfg.ELEM == 8 else 0, tag="rows")]
    tiles = -(-n1 // cfg.TX)
    slots = device_props.sm_count * cfg.blocks_per_sm
    floors = (max(4 * cfg.U, 64), 2 * cfg.U)

    def geometry(a0_begin, a0_end):
        n_units = a0_end - a0_begin
        chunks = plan_chunks_adaptive(n_units, tiles, slots, cfg.W1, floors, device_props.sm_count)
        chunk_len = -(-n_units // chunks)
        chunks = -(-n_units // chunk_len)
        launch = Launch((tiles, 1, chunks), (cfg.threads, 1, 1), cfg.smem_bytes)
        launch.extra = (chunk_len,)
        return launch

    notes = dict(cfg.describe(), rank=2, strategy="rows", mode="regwin", embedded_plane=False, static_smem=0)
    return CudaPlan(program, domain, 'stream', source, SINGLE_FUNCTION, options, geometry,
                    tensor_maps=specs, notes={'config': notes})
