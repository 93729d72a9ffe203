"""Strategy 'temporal': two sweeps of an iterated rank-3 stencil fused into one pass over HBM.

For ``stencil.iterate(grid, n)`` (Jacobi / Laplacian style iteration, SURVEY.md section 7 step 6;
the idea of the reference's notebook ``ipython_notebooks/fusing_stencils``: keep the intermediate
result of the first application on chip).  Per launch the kernel computes out = S(S(in)) with
8 B of HBM traffic per point for TWO sweeps.

Pipeline (extends ``codegen_stream``): the producer warp streams input planes of a tile through
the TMA ring; consumer threads run sweep 1 from the rotating input register window and obtain the
intermediate ("mid") plane z-1 for their own BY x 4 block.  Own mid values stay in a second
register window (they are the z-1 / z / z+1 centre taps of sweep 2); the block is also written
to one of three shared-memory mid buffers, a named barrier over the consumer warps follows, and
sweep 2 reads its in-plane neighbours of the previous mid plane from shared memory and stores
output plane z-2.  Tiles overlap: a CTA computes mid on MY x 128 points and outputs the interior
(MY - 2*hy) x 120 of it, so no CTA needs another CTA's intermediate values.

Boundary semantics are those of two consecutive sweeps: mid is subject to the same boundary
handling as any sweep result ('zero'/'copy': halo selected at production; 'clamp': sweep 2 reads
mid through the same clamped row offsets / column cascade as sweep 1 reads the input, and a
uniform select substitutes the nearest in-grid mid plane along axis 0).  The arithmetic of both
sweeps is the unchanged per-point program: results are bit-identical to two separate sweeps.

Eligibility: rank 3, float32, exactly one grid argument, taps off the centre column only at
dz = 0, |dz| <= 1, column radius <= 4, no data-indexed tables.
"""
import math
import os

from ...stencil_exception import StencilException
from .. import point_program as pp
from .codegen_common import PROLOGUE, emit_statements
from .codegen_stream import (PTX_HELPERS, TensorMapSpec, Footprint, _Emitter, _group_loads, _tap3,
                             _grid_args, table_plan, plan_chunks, _const_index)

FUNCTION = "stencil_temporal2"


def eligible(program, domain):
    if program.dim != 3 or program.out_ctype != pp.FLOAT:
        return False, "rank 3 float32 only"
    grids = _grid_args(program)
    if len(grids) != 1 or grids[0].index != 0 or grids[0].ctype != pp.FLOAT:
        return False, "exactly one float32 grid argument"
    if domain.row_pitch % 4 or program.shape[-1] < 256 or program.shape[1] < 16 or program.shape[0] < 8:
        return False, "grid too small or rows not 16-byte multiples"
    if domain.is_slab and domain.slab_ghost < 2 * program.ghost_depth[0]:
        return False, "slab carries fewer than two ghost depths of planes"
    data_tables = table_plan(program, False)[0]
    if data_tables:
        return False, "data-indexed tables"
    if any(isinstance(s, pp.LocalAssign) for s in program.statements):
        return False, "local temporaries"
    ages = set()
    for tap in program.taps():
        dz, dy, dx = _tap3(tap.offset)
        ages.add(dz)
        if dz != 0 and (dy or dx):
            return False, "taps off the centre column outside the middle plane"
        if abs(dz) > 1 or abs(dx) > 4:
            return False, "tap radius"
    if ages != {-1, 0, 1}:
        return False, "needs taps at dz = -1, 0, +1"
    return True, "star-in-z rank-3 stencil"


def pair_eligible(first, first_domain, second, second_domain):
    """Can ``second(first(grid))`` -- two DIFFERENT stencils -- run as one fused launch?  Each must
    be eligible on its own, take the grid as its only argument (no tables: the two parameter lists
    are not merged), and agree on shape and boundary mode."""
    for program, domain in ((first, first_domain), (second, second_domain)):
        ok, why = eligible(program, domain)
        if not ok:
            return False, why
        if len(program.args) != 1:
            return False, "stencils with table arguments are not fused with other stencils"
        if domain.is_slab:
            return False, "slab domains"
    if first.shape != second.shape or first.boundary != second.boundary:
        return False, "shape or boundary mode differ"
    return True, "pair of star-in-z stencils"


def _extents(program, second=None):
    lo, hi = program.tap_extent()
    hx, hy = max(-lo[2], hi[2]), max(-lo[1], hi[1])
    if second is not None:
        lo, hi = second.tap_extent()
        hx, hy = max(hx, -lo[2], hi[2]), max(hy, -lo[1], hi[1])
    return hx, hy


def worthwhile(program, cfg):
    """Does the fused kernel beat two single sweeps on this grid?  It is issue-bound, so its time
    follows the points it computes, partial tiles included: measured on B200, 1024^2 planes
    (83 % of the computed mid points useful) run 1.28x faster than two sweeps, 256^2 planes
    (56 %) 0.86x.  Break-even is near 65 %."""
    n0, n1, n2 = program.shape
    tiles_x = -(-n2 // cfg.OUT_COLS)
    tiles_y = -(-n1 // cfg.OUT_ROWS)
    useful = n1 * n2 / float(tiles_y * cfg.MY * tiles_x * 128)
    return useful >= 0.68


def _tag(v):
    return "m{}".format(-v) if isinstance(v, int) and v < 0 else str(v)


class TemporalConfig(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def describe(self):
        keys = ("MY", "BY", "NCW", "S", "HX", "HY", "OUT_ROWS", "OUT_COLS", "blocks_per_sm", "smem_bytes")
        return dict({k: getattr(self, k) for k in keys}, strategy="temporal2")


def choose_config(program, device_props, second=None):
    # a pair of different stencils shares one tile geometry: halos sized for the wider of the two
    hx, hy = _extents(program, second)
    PADX = 4 if hx else 0
    override = {}
    for item in filter(None, os.environ.get("STENCIL_CUDA_TEMPORAL_TILE", "").split(",")):
        key, _, value = item.partition("=")
        override[key.strip()] = int(value)
    best = None
    # measured on B200 (Jacobi-3D 512x1024x1024, GStencil/s): (16,2,8,p2) 977, (32,2,16,p1) 966,
    # (32,2,16,p2) 926-954; register blocks of 4 rows spill under the occupancy bound (355-377)
    for MY, BY, NCW, prefetch in ((16, 2, 8, 4), (16, 2, 8, 2), (32, 2, 16, 1), (32, 2, 16, 2), (8, 1, 8, 2)):
        if override:
            MY, BY, NCW = override.get("MY", MY), override.get("BY", BY), override.get("NCW", NCW)
            prefetch = override.get("prefetch", prefetch)
        if MY != BY * NCW or MY <= 2 * hy:
            continue
        S = 2 + prefetch
        pitch = 128 + 2 * PADX
        rows = MY + 2 * hy
        stage_floats = (rows * pitch * 4 + 127) // 128 * 32
        mid_floats = MY * 128
        pad_floats = max(1, hy) * 128                 # rows read above / below the mid buffers
        smem = S * stage_floats * 4 + (3 * mid_floats + 2 * pad_floats) * 4 + 2 * S * 8 + 128
        threads = (NCW + 1) * 32
        if smem > device_props.smem_per_block_optin:
            continue
        blocks = min(device_props.smem_per_sm // (smem + 1024), device_props.max_threads_per_sm // threads, 4)
        blocks = min(blocks, max(1, device_props.regs_per_sm // (threads * 104)))
        if blocks < 1:
            continue
        out_rows, out_cols = MY - 2 * hy, 128 - (8 if hx else 0)
        # traffic per output point of the pair of sweeps, and redundant sweep-1 work
        traffic = rows * (pitch + 8) / float(out_rows * out_cols) + 1.0
        redundancy = MY * 128 / float(out_rows * out_cols)
        score = traffic * (1.0 + 0.15 * (redundancy - 1.0)) * (1.0 if blocks * NCW >= 16 else 1.08)
        config = TemporalConfig(MY=MY, BY=BY, NCW=NCW, S=S, HX=hx, HY=hy, PADX=PADX, PITCH=pitch, ROWS=rows,
                                STAGE_FLOATS=stage_floats, STAGE_BYTES=rows * pitch * 4, MID_FLOATS=mid_floats,
                                PAD_FLOATS=pad_floats,
                                OUT_ROWS=out_rows, OUT_COLS=out_cols, smem_bytes=smem, threads=threads,
                                blocks_per_sm=blocks, score=score)
        if best is None:
            best = config                      # candidates are in measured order: first that fits
        if override:
            break
    if best is None:
        raise StencilException("no temporal tile fits in shared memory")
    return best


STEP_HELPERS = r'''
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
'''


def generate(program, domain, cfg, function=FUNCTION, mirror=None, prologue=True, second=None,
             second_domain=None, step_kernel=False):
    """CUDA source of the fused kernel.  ``mirror``: also store results through ``out_peer`` (slab
    boundary launches); defaults to ``domain.is_slab``.  ``second``: the program of the second
    sweep when it is another stencil than the first (``fuse(a, b)``); its halo (``zero``/``copy``)
    follows ``second_domain``, i.e. its own ghost depth.

    ``step_kernel``: ONE launch per exchange step of a slab (``multi_gpu.SlabStencil``).  The
    grid's first z-slices are the boundary planes next to each neighbouring slab -- scheduled
    first; they wait (the producer's elected thread, ``ld.acquire.sys``) until the neighbour's flag
    says its previous step's boundary planes are in this slab's ghost planes, store their results
    locally and into the neighbour's ghost planes (peer memory over NVLink), and the last of them
    to finish publishes this slab's flag in the neighbours' memory (``st.release.sys`` after a
    system-scope fence).  The remaining z-slices are the interior chunks, which never touch ghost
    planes: they hide the exchange."""
    mirror = domain.is_slab if mirror is None else mirror
    if step_kernel:
        mirror = True
    second = program if second is None else second
    second_domain = domain if second_domain is None else second_domain
    shape = program.shape
    N0, N1, N2 = shape
    MY, BY, NCW, S = cfg.MY, cfg.BY, cfg.NCW, cfg.S
    HX, HY, PADX, PITCH = cfg.HX, cfg.HY, cfg.PADX, cfg.PITCH
    SF, MF = cfg.STAGE_FLOATS, cfg.MID_FLOATS
    MP = 128                                  # mid buffer pitch
    boundary = program.boundary
    clamp, copy = boundary == 'clamp', boundary == 'copy'
    c_lo, c_hi = domain.clamp_lo, domain.clamp_hi
    i_lo, i_hi = domain.interior_lo, domain.interior_hi
    j_lo, j_hi = second_domain.interior_lo, second_domain.interior_hi      # second sweep's interior
    two_masks = (tuple(i_lo), tuple(i_hi)) != (tuple(j_lo), tuple(j_hi))
    fp = Footprint(program, BY, copy)
    _, const_refs, _, _ = table_plan(program, False)
    clamp_rows = clamp and HY > 0
    clamp_cols = clamp and HX > 0
    W = 3

    def in_slot(phase, age, key):
        ages = fp.ages[key]
        if max(ages) == min(ages):
            return 0
        return (phase + age - 1) % W

    def in_name(phase, age, row, col):
        return "i{}_{}_{}".format(in_slot(phase, age, (0, row, col)), _tag(row), _tag(col))

    def mid_own(phase, age, b, j):
        """own mid value of the plane with `age` relative to the OUTPUT plane (computed at unit-1+age)"""
        return "m{}_{}_{}".format((phase + age - 1) % W, b, j)

    def table_text(ref, index_text):
        return "tc{}_{}".format(ref.arg, _const_index(ref.index))

    e = _Emitter()
    if prologue:
        e.lines.append(PROLOGUE)
        e.lines.append(PTX_HELPERS.replace("@STOP@", os.environ.get("STENCIL_CUDA_STORE_OP", ".cs")))
    e("// temporal strategy: two fused sweeps | grid {} boundary {} | mid tile {}x128, output {}x{}, "
      "BY={} ring S={}".format(shape, boundary, MY, cfg.OUT_ROWS, cfg.OUT_COLS, BY, S))
    params = ["const {}* __restrict__ a{}".format(arg.ctype, arg.index) for arg in program.args]
    if step_kernel:
        e.lines.append(STEP_HELPERS)
        params += ["const __grid_constant__ TensorMap tm0", "float* __restrict__ out",
                   "int a0_begin", "int a0_end", "int chunk_len",
                   "float* __restrict__ peer_up", "i64 shift_up", "float* __restrict__ peer_dn", "i64 shift_dn",
                   "int top_begin", "int top_end", "int bot_begin", "int bot_end",
                   "u32* sync", "u32* flag_up", "u32* flag_dn", "u32 step"]
    else:
        params += ["const __grid_constant__ TensorMap tm0", "float* __restrict__ out",
                   "float* __restrict__ out_peer", "i64 peer_shift", "int a0_begin", "int a0_end", "int chunk_len"]
    e('extern "C" __global__ void __launch_bounds__({}, {})'.format(cfg.threads, cfg.blocks_per_sm))
    e("{}({})".format(function, ", ".join(params)))
    e.open("")
    e("extern __shared__ __align__(128) unsigned char smem_raw[];")
    e("float* const ring = reinterpret_cast<float*>(smem_raw);")
    e("float* const mid = ring + {} + {};        // three mid buffers, hy pad rows on either side".format(
        S * SF, cfg.PAD_FLOATS))
    e("u64* const bar_full = reinterpret_cast<u64*>(smem_raw + {});".format((S * SF + 3 * MF + 2 * cfg.PAD_FLOATS) * 4))
    e("u64* const bar_empty = bar_full + {};".format(S))
    e("const int tid = threadIdx.x;")
    e("const int warp = tid >> 5;")
    e("const int lane = tid & 31;")
    e("const int x0 = blockIdx.x * {} - {};     // origin of the mid tile (outputs are its interior)".format(
        cfg.OUT_COLS, PADX))
    e("const int y0 = blockIdx.y * {} - {};".format(cfg.OUT_ROWS, HY))
    n_edge = int(domain.open_lo) + int(domain.open_hi) if step_kernel else 0
    if step_kernel:
        e("// z-slices: the chunk(s) at the slab's open end(s) first ({} neighbour{}), then the middle chunks".format(
            n_edge, "" if n_edge == 1 else "s"))
        e("int u_begin, u_end, mir_lo = 0, mir_hi = 0;      // [mir_lo, mir_hi): planes that also go to the neighbour")
        e("float* out_peer = nullptr; i64 peer_shift = 0; const u32* wait_flag = nullptr;")
        e("const bool edge = blockIdx.z < {};".format(n_edge))
        first = True
        if domain.open_lo:
            e.open("if (blockIdx.z == 0)")
            e("u_begin = top_begin; u_end = top_end; out_peer = peer_up; peer_shift = shift_up; wait_flag = sync;")
            e("mir_lo = top_begin; mir_hi = top_begin + {};".format(domain.slab_ghost))
            e.close()
            first = False
        if domain.open_hi:
            e.open("{}if (blockIdx.z == {})".format("" if first else "else ", 0 if first else 1))
            e("u_begin = bot_begin; u_end = bot_end; out_peer = peer_dn; peer_shift = shift_dn; wait_flag = sync + 1;")
            e("mir_lo = bot_end - {}; mir_hi = bot_end;".format(domain.slab_ghost))
            e.close()
            first = False
        e.open("{}".format("" if first else "else"))
        e("u_begin = a0_begin + ((int)blockIdx.z - {}) * chunk_len;".format(n_edge))
        e("u_end = min(u_begin + chunk_len, a0_end);")
        e.close()
    else:
        e("const int u_begin = a0_begin + blockIdx.z * chunk_len;")
        e("const int u_end = min(u_begin + chunk_len, a0_end);")
    e("if (u_begin >= u_end) return;")
    e("const int p_first = u_begin - 2;             // axis-0 index of streamed input plane 0")
    e("const int n_units = (u_end - u_begin) + 4;")
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
    if step_kernel:
        e("// boundary planes read ghost planes: the neighbour's previous step must have delivered them")
        e("// (and have finished reading the ghost planes this step overwrites in ITS memory)")
        e.open("if (edge && lane == 0)")
        e("while (ld_acquire_sys(wait_flag) < step) { }")
        e.close()
        e("__syncwarp();")
    e.open("for (int k = 0; k < n_units; ++k)")
    e("const int s = k % {};".format(S))
    e("if (k >= {S}) mbar_wait(&bar_empty[s], ((k / {S}) - 1) & 1);".format(S=S))
    e.open("if (lane == 0)")
    e("mbar_expect_tx(&bar_full[s], {});".format(cfg.STAGE_BYTES))
    e("tma_load_3d(ring + s * {sf}, &tm0, &bar_full[s], x0 - {px}, y0 - {hy}, SC_CLAMP(p_first + k, {lo}, {hi}));".format(
        sf=SF, px=PADX, hy=HY, lo=c_lo[0], hi=c_hi[0]))
    e.close()
    e.close()
    e.lines.append("#ifdef SC_PDL")
    e("mbar_wait(&bar_full[(n_units - 1) % {S}], ((n_units - 1) / {S}) & 1);    // all loads have landed".format(S=S))
    e("pdl_launch_dependents();")
    e.lines.append("#endif")
    e("return;")
    e.close()
    e()
    # ------------------------------------------------------------------ consumers
    # The loop body is emitted once per phase of lcm(W, S) phases: the register-window slot, the
    # ring stage and the mid buffer of every access are then compile-time constants (immediate
    # offsets of LDS/STS, no modulo arithmetic); only the mbarrier parity is carried.
    P = W * S // math.gcd(W, S)
    use_shfl = os.environ.get("STENCIL_CUDA_TEMPORAL_SHFL", "1") not in ("0", "")
    in_shfl = os.environ.get("STENCIL_CUDA_TEMPORAL_INSHFL", "1") not in ("0", "")
    own_regs = os.environ.get("STENCIL_CUDA_TEMPORAL_OWNREGS", "1") not in ("0", "")
    own_in_block = own_regs or not (clamp and HY > 0)       # sweep 2 reads own-block rows from registers
    e("const int tr = warp;")
    e("const int gy = y0 + tr * {};               // first row of this thread's block".format(BY))
    e("const int gx = x0 + lane * 4;             // first column of this thread's block")
    e("const int sbase_in = ({} + tr * {}) * {} + {} + lane * 4;".format(HY, BY, PITCH, PADX))
    e("const int sbase_mid = tr * {} * {} + lane * 4;".format(BY, MP))
    rows_in = sorted({k[1] for k in fp.ages})
    rows_mid = sorted({b + _tap3(tap.offset)[1] for tap in second.taps() for b in range(BY)
                       if _tap3(tap.offset)[0] == 0} | set(range(BY)))
    # per-row base pointers (the clamped row offset of 'clamp' folded in): loads are pointer[immediate]
    for row in rows_in:
        delta = "(min(max(gy + ({r}), 0), {n}) - gy)".format(r=row, n=N1 - 1) if clamp_rows else "({})".format(row)
        e("const float* const pI{t} = ring + sbase_in + {d} * {p};".format(t=_tag(row), d=delta, p=PITCH))
    for row in rows_mid:
        delta = "(min(max(gy + ({r}), 0), {n}) - gy)".format(r=row, n=N1 - 1) if clamp_rows else "({})".format(row)
        e("const float* const pM{t} = mid + sbase_mid + {d} * {p};".format(t=_tag(row), d=delta, p=MP))
    e("float* const mid_own_rows = mid + sbase_mid;")
    if clamp_cols:
        e("const bool x_edge = x0 <= 0 || x0 + {} > {};".format(128 + HX, N2))
    if clamp_rows and own_regs and BY > 1:
        e("const bool y_edge = y0 < 0 || y0 + {} > {};".format(MY, N1))
        for b in range(BY):
            e("const int cr{b} = min(max(gy + {b}, 0), {n}) - gy;   // block row that row {b} reads as under 'clamp'".format(
                b=b, n=N1 - 1))
    if not clamp:
        e("u32 halo_mask = 0;                   // bit (b*4+j): the point lies in the in-plane halo")
        for b in range(BY):
            for j in range(4):
                tests = []
                if i_lo[1] > 0:
                    tests.append("gy + {} < {}".format(b, i_lo[1]))
                if i_hi[1] < N1:
                    tests.append("gy + {} >= {}".format(b, i_hi[1]))
                if i_lo[2] > 0:
                    tests.append("gx + {} < {}".format(j, i_lo[2]))
                if i_hi[2] < N2:
                    tests.append("gx + {} >= {}".format(j, i_hi[2]))
                if tests:
                    e("if ({}) halo_mask |= {}u;".format(" || ".join(tests), 1 << (b * 4 + j)))
        if two_masks:
            e("u32 halo_mask2 = 0;                  // the same for the second stencil's ghost depth")
            for b in range(BY):
                for j in range(4):
                    tests = []
                    if j_lo[1] > 0:
                        tests.append("gy + {} < {}".format(b, j_lo[1]))
                    if j_hi[1] < N1:
                        tests.append("gy + {} >= {}".format(b, j_hi[1]))
                    if j_lo[2] > 0:
                        tests.append("gx + {} < {}".format(j, j_lo[2]))
                    if j_hi[2] < N2:
                        tests.append("gx + {} >= {}".format(j, j_hi[2]))
                    if tests:
                        e("if ({}) halo_mask2 |= {}u;".format(" || ".join(tests), 1 << (b * 4 + j)))
    # which of this thread's outputs are inside the valid (interior) part of the mid tile and the grid
    lane_ok = "lane >= 1 && lane <= 30 && " if HX else ""
    e("const bool col_ok = {}gx >= 0 && gx < {};".format(lane_ok, N2))
    for b in range(BY):
        e("const bool store_ok{b} = col_ok && tr * {by} + {b} >= {hy} && tr * {by} + {b} < {hi} && "
          "gy + {b} >= 0 && gy + {b} < {n1};".format(b=b, by=BY, hy=HY, hi=MY - HY, n1=N1))
    e("float* dst = out + ((i64)u_begin * {} + gy) * {} + gx;   // this thread's block in the next output plane".format(
        N1, domain.row_pitch))
    if mirror:
        e("const i64 mirror = out_peer ? (out_peer - out) + peer_shift : 0;")
    for (arg, index), ctype in sorted(const_refs.items()):
        e("const {} tc{}_{} = a{}[{}];".format(ctype, arg, index, arg, index))
    names = []
    for (g, row, col), ages in sorted(fp.ages.items()):
        slots = W if max(ages) > min(ages) else 1
        names += ["i{}_{}_{}".format(slot, _tag(row), _tag(col)) for slot in range(slots)]
    names += ["m{}_{}_{}".format(slot, b, j) for slot in range(W) for b in range(BY) for j in range(4)]
    for first in range(0, len(names), 8):
        e("float {};".format(", ".join(n + " = 0.0f" for n in names[first:first + 8])))
    counter = [0]

    def column_fix(names_by_col, in_quad=False):
        """``in_quad``: rows that end inside the quad (ragged grids): start the cascade at column 1"""
        for c in sorted((c for c in names_by_col if c < 0), reverse=True):
            e("if (gx + ({}) < 0) {} = {};".format(c, names_by_col[c], names_by_col[c + 1]))
        for c in sorted(c for c in names_by_col if c > (0 if in_quad else 3)):
            e("if (gx + {} > {}) {} = {};".format(c, N2 - 1, names_by_col[c], names_by_col[c - 1]))

    def load_rows(prefix, offset, by_row, name_of_col, declare):
        """vectorised shared-memory loads of {row: cols} through the per-row pointers ``prefix<row>``
        at the compile-time offset ``offset`` (ring stage / mid buffer)"""
        for row in sorted(by_row):
            pointer = "{}{}".format(prefix, _tag(row))
            for first, width, used in _group_loads(by_row[row]):
                if width == 1:
                    e("{}{} = {}[{}];".format(declare, name_of_col(row, used[0]), pointer, offset + first))
                    continue
                counter[0] += 1
                tmp = "q{}".format(counter[0])
                vec = "float4" if width == 4 else "float2"
                e("const {v} {t} = *reinterpret_cast<const {v}*>({p} + {o});".format(
                    v=vec, t=tmp, p=pointer, o=offset + first))
                e(" ".join("{}{} = {}.{};".format(declare, name_of_col(row, c), tmp, "xyzw"[c - first]) for c in used))

    def emit_loop(mirror):
        """the consumers' plane loop; ``mirror``: results also go to the neighbour's ghost planes"""
        e("int unit = 0;                        // newest streamed input plane, relative to p_first")
        e("u32 par = 0;                         // parity of the ring's full barriers in this pass over the phases")
        e.open("while (true)")
        for phase in range(P):
            stage = phase % S
            e("// ---- phase {} (window slot {}, ring stage {}, mid buffer {})".format(phase, phase % W, stage, phase % 3))
            e("if (unit >= n_units) break;")
            e.open("")
            fills = phase // S                       # how often this stage was filled before, within this pass
            parity = "par" if fills % 2 == 0 else "(par ^ 1u)"
            if (P // S) % 2 == 0:
                parity = str(fills % 2)
            e("mbar_wait(&bar_full[{}], {});".format(stage, parity))
            # ---------------- sweep-1 input window (same scheme as codegen_stream, regwin)
            reads = {}
            for key, ages in fp.ages.items():
                reads.setdefault(max(ages), []).append(key)
            for age in sorted(reads):
                back = 1 - age
                by_row, lane_edge = {}, {}
                for (_, row, col) in reads[age]:
                    source = fp.ages.get((0, row, col % 4))
                    if in_shfl and (col < 0 or col > 3) and -4 <= col < 8 and source and min(source) <= age <= max(source):
                        lane_edge.setdefault(row, set()).add(col)
                    else:
                        by_row.setdefault(row, set()).add(col)
                load_rows("pI", ((phase - back) % S) * SF, by_row,
                          lambda row, col, age=age: in_name(phase, age, row, col), "")
                # columns next to the quad: the neighbouring lane holds them in its window registers.
                # Only the warp's first / last lane reads shared memory (the tile's column halo): one
                # conflict-free wavefront instead of a four-way bank conflict per value.
                for row in sorted(lane_edge):
                    offset = ((phase - back) % S) * SF
                    for col in sorted(lane_edge[row]):
                        left = col < 0
                        name = in_name(phase, age, row, col)
                        e("{n} = __shfl_{d}_sync(0xffffffffu, {src}, 1);".format(
                            n=name, d="up" if left else "down", src=in_name(phase, age, row, col % 4)))
                        e("if (lane == {l}) {n} = pI{t}[{o}];".format(l=0 if left else 31, n=name, t=_tag(row), o=offset + col))
                for row, cols in lane_edge.items():
                    by_row.setdefault(row, set()).update(cols)         # for the column cascade below
                if clamp_cols and any(c < 0 or c > 3 for cols in by_row.values() for c in cols):
                    e.open("if (x_edge)")
                    for row in sorted(by_row):
                        outer = [c for c in by_row[row] if c < 0 or c > 3]
                        if outer:
                            span = range(min(min(outer), 0), max(max(outer), 3) + 1)
                            column_fix({c: in_name(phase, age, row, c) for c in span if (0, row, c) in fp.ages},
                                       in_quad=fp.ragged)
                    e.close()
            release_back = 1 - min(max(s) for s in fp.ages.values())
            e("__syncwarp();")
            e("if (lane == 0 && unit >= {b}) mbar_arrive(&bar_empty[{s}]);".format(b=release_back, s=(phase - release_back) % S))

            # ---------------- sweep 1: mid plane zm = p_first + unit - 1
            e.open("if (unit >= 2)")
            e("const int zm = p_first + unit - 1;")
            macc = [[mid_own(phase, 1, b, j) for j in range(4)] for b in range(BY)]     # straight into the window slot
            for b in range(BY):
                e(" ".join("{} = 0.0f;".format(n) for n in macc[b]))
            for statement in program.statements:
                for b in range(BY):
                    for j in range(4):
                        def tap_text(tap, b=b, j=j):
                            dz, dy, dx = _tap3(tap.offset)
                            return in_name(phase, dz, b + dy, j + dx)
                        single = pp.PointProgram(3, shape, program.out_ctype, program.args, [statement],
                                                 program.ghost_depth, boundary)
                        e(emit_statements(single, tap_text, table_text, acc=macc[b][j], indent="")[0])
            if not clamp:
                ztests = []
                if i_lo[0] > 0:
                    ztests.append("zm < {}".format(i_lo[0]))
                if i_hi[0] < N0:
                    ztests.append("zm >= {}".format(i_hi[0]))
                e("const bool zm_halo = {};".format(" || ".join(ztests) if ztests else "false"))
                e.open("if (zm_halo || halo_mask)")
                e('asm volatile("" ::: "memory");     // keeps this rare case a branch, not selects for everyone')
                for b in range(BY):
                    for j in range(4):
                        value = in_name(phase, 0, b, j) if copy else "0.0f"
                        e("if (zm_halo || (halo_mask & {}u)) {} = {};".format(1 << (b * 4 + j), macc[b][j], value))
                e.close()
            if clamp_cols and fp.ragged:
                # the intermediate grid's columns past the last one read as the last one
                e.open("if (x_edge)")
                for b in range(BY):
                    column_fix({j: macc[b][j] for j in range(4)}, in_quad=True)
                e.close()
            if clamp_rows and own_regs and BY > 1:
                # rows of the thread's own block that lie outside the grid read as the clamped row: where
                # that row is in the same block, copy it now (a uniform branch, taken by edge tiles only),
                # so that sweep 2 can take own-block rows from registers instead of shared memory
                e.open("if (y_edge)")
                e('asm volatile("" ::: "memory");')
                for b in range(BY):
                    for c in range(BY):
                        if c != b:
                            e("if (cr{} == {}) {{ {} }}".format(b, c, " ".join(
                                "{} = {};".format(mid_own(phase, 1, b, j), mid_own(phase, 1, c, j)) for j in range(4))))
                e.close()
            if clamp:
                # axis-0 clamp of the intermediate grid, applied where a plane is produced (a uniform,
                # rarely taken branch) instead of by two selects per point where it is used: the plane
                # past the last one reads as the last one, the plane before the first one as the first
                e.open("if (zm <= {} || zm > {})".format(c_lo[0], c_hi[0]))
                e('asm volatile("" ::: "memory");          // keeps this a branch (not 2 selects per point)')
                for b in range(BY):
                    for j in range(4):
                        e("if (zm > {hi}) {new} = {prev}; else if (zm == {lo}) {prev} = {new};".format(
                            hi=c_hi[0], lo=c_lo[0], new=mid_own(phase, 1, b, j), prev=mid_own(phase, 0, b, j)))
                e.close()
            for b in range(BY):
                e("*reinterpret_cast<float4*>(mid_own_rows + {}) = make_float4({});".format(
                    (phase % 3) * MF + b * MP, ", ".join(macc[b])))
            e('asm volatile("bar.sync 1, {};" ::: "memory");'.format(NCW * 32))

            # ---------------- sweep 2: output plane zo = p_first + unit - 2
            e.open("if (unit >= 4)")
            mid_offset = ((phase + 2) % 3) * MF                    # buffer of mid plane zo
            # in-plane taps of sweep 2.  A tap in the point's own row and own quad is the thread's own
            # mid register.  Everything else is read from the mid buffer; under 'clamp' that includes
            # other rows of the thread's own block, because a block row outside the grid must read as
            # the clamped row (the buffer read goes through the clamped row pointer, a register would
            # not).  Columns next to the quad in the point's own row come from the neighbouring lanes'
            # registers by warp shuffle (no bank-conflicting scalar shared-memory loads).
            by_row = {}
            shuffled = {}                     # (row, col) -> (source register of the neighbour lane, lane delta)
            for tap in second.taps():
                dz, dy, dx = _tap3(tap.offset)
                if dz != 0:
                    continue
                for b in range(BY):
                    for j in range(4):
                        row, col = b + dy, j + dx
                        own = 0 <= col < 4 and (dy == 0 if (clamp and not own_in_block) else 0 <= row < BY)
                        if own:
                            continue
                        if use_shfl and dy == 0 and -4 <= col < 8:
                            shuffled[(row, col)] = (mid_own(phase, 0, row, col % 4), -1 if col < 0 else 1)
                        else:
                            by_row.setdefault(row, set()).add(col)
            if clamp_cols:
                for row, cols in by_row.items():       # the select cascade needs every column up to the quad
                    if min(cols) < 0:
                        cols.update(range(min(cols), 0))
                    if max(cols) > 3:
                        cols.update(range(4, max(cols) + 1))
            for key in [k for k in shuffled if k[1] in by_row.get(k[0], ())]:
                del shuffled[key]                      # already read from the buffer for a tap in another row
            if clamp_cols:
                for (row, col) in list(shuffled):
                    span = range(col, 0) if col < 0 else range(4, col + 1)
                    for c in span:
                        if (row, c) not in shuffled and c not in by_row.get(row, ()):
                            shuffled[(row, c)] = (mid_own(phase, 0, row, c % 4), -1 if c < 0 else 1)

            def nb_name(row, col, same_row=False):
                own = 0 <= col < 4 and (same_row if (clamp and not own_in_block) else 0 <= row < BY)
                if own:
                    return mid_own(phase, 0, row, col)
                return "n_{}_{}".format(_tag(row), _tag(col))
            load_rows("pM", mid_offset, by_row, nb_name, "float ")
            for (row, col), (source, delta) in sorted(shuffled.items()):
                e("float n_{}_{} = __shfl_{}_sync(0xffffffffu, {}, 1);".format(
                    _tag(row), _tag(col), "up" if delta < 0 else "down", source))
            outer_rows = {}
            for row, cols in by_row.items():
                outer_rows.setdefault(row, set()).update(c for c in cols if c < 0 or c > 3)
            for (row, col) in shuffled:
                outer_rows.setdefault(row, set()).add(col)
            if clamp_cols and any(outer_rows.values()):
                e.open("if (x_edge)")
                for row in sorted(outer_rows):
                    outer = sorted(outer_rows[row])
                    if outer:
                        # the cascade starts from the thread's own quad: columns 0 and 3 of this row,
                        # loaded from the buffer if the row is not the point's own, else the register
                        names = {c: nb_name(row, c) for c in outer}
                        for c in (0, 3):
                            names[c] = ("n_{}_{}".format(_tag(row), c) if c in by_row.get(row, ())
                                        else mid_own(phase, 0, row, c))
                        column_fix(names)
                e.close()
            acc = [["acc{}_{}".format(b, j) for j in range(4)] for b in range(BY)]
            for b in range(BY):
                e("float {};".format(", ".join(n + " = 0.0f" for n in acc[b])))
            for statement in second.statements:
                for b in range(BY):
                    for j in range(4):
                        def tap_text(tap, b=b, j=j):
                            dz, dy, dx = _tap3(tap.offset)
                            if dz == 0:
                                return nb_name(b + dy, j + dx, same_row=(dy == 0))
                            return mid_own(phase, dz, b, j)
                        single = pp.PointProgram(3, shape, second.out_ctype, second.args, [statement],
                                                 second.ghost_depth, boundary)
                        e(emit_statements(single, tap_text, table_text, acc=acc[b][j], indent="")[0])
            if not clamp:
                e("const int zo = p_first + unit - 2;")
                mask2 = "halo_mask2" if two_masks else "halo_mask"
                ztests = []
                if j_lo[0] > 0:
                    ztests.append("zo < {}".format(j_lo[0]))
                if j_hi[0] < N0:
                    ztests.append("zo >= {}".format(j_hi[0]))
                e("const bool zo_halo = {};".format(" || ".join(ztests) if ztests else "false"))
                e.open("if (zo_halo || {})".format(mask2))
                e('asm volatile("" ::: "memory");     // keeps this rare case a branch, not selects for everyone')
                for b in range(BY):
                    for j in range(4):
                        # 'copy': the halo of sweep 2 copies the halo of mid, which copied the input
                        value = mid_own(phase, 0, b, j) if copy else "0.0f"
                        e("if (zo_halo || ({} & {}u)) {} = {};".format(mask2, 1 << (b * 4 + j), acc[b][j], value))
                e.close()
            if mirror and step_kernel:
                e("const int zs = p_first + unit - 2;")
                e("const bool to_peer = zs >= mir_lo && zs < mir_hi;")
            for b in range(BY):
                e.open("if (store_ok{})".format(b))
                e("st_global_v4(dst + {}, {});".format(b * domain.row_pitch, ", ".join(acc[b])))
                if mirror:                  # slab boundary launches also fill the neighbour's ghost planes
                    e("if ({}) st_global_v4(dst + {} + mirror, {});".format(
                        "to_peer" if step_kernel else "out_peer", b * domain.row_pitch, ", ".join(acc[b])))
                e.close()
            e("dst += {};".format(N1 * domain.row_pitch))
            e.close()   # unit >= 4
            e.close()   # unit >= 2
            e("++unit;")
            e.close()
        if (P // S) % 2 == 1:
            e("par ^= 1u;")
        e.close()   # while

    if step_kernel:
        e.open("if (edge)")
        emit_loop(True)
        # every boundary CTA of this step has stored its planes (here and in the neighbour's ghost
        # planes): the last one to arrive publishes the step number in the neighbours' memory
        e('asm volatile("bar.sync 1, {};" ::: "memory");'.format(NCW * 32))
        e.open("if (tid == 0)")
        e("__threadfence_system();")
        e("const u32 arrived = atomicAdd(sync + 2, 1u);")
        e.open("if (arrived == gridDim.x * gridDim.y * {}u - 1u)".format(n_edge))
        e("sync[2] = 0;                       // for the next step (which starts after this grid has completed)")
        e("__threadfence_system();")
        e("if (flag_up) st_release_sys(flag_up, step + 1u);")
        e("if (flag_dn) st_release_sys(flag_dn, step + 1u);")
        e.close()
        e.close()
        e.close()
        e.open("else")
        emit_loop(False)
        e.close()
    else:
        emit_loop(mirror)
    e.close()   # kernel
    return e.text()


def make_plan(program, domain, device_props, options, second=None, second_domain=None):
    """CudaPlan-like record for the fused two-sweep kernel (launched by ``iterate``; with
    ``second`` by ``fuse(first, second)``)."""
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
    if domain.is_slab:
        # interior launches of a slab never mirror: a second entry point without that code
        # (the kernel is issue bound; the unused mirror stores cost about 6 %)
        source += generate(program, domain, cfg, function=FUNCTION + "_interior", mirror=False, prologue=False)
        if second is None:
            # one launch per exchange step: boundary planes first, then the interior (multi_gpu.py)
            source += generate(program, domain, cfg, function=FUNCTION + "_step", prologue=False, step_kernel=True)
    n0, n1, n2 = program.shape
    row_pitch = domain.row_pitch
    specs = [TensorMapSpec(0, (n2, n1, n0), (row_pitch * 4, n1 * row_pitch * 4), (cfg.PITCH, cfg.ROWS, 1),
                           tag="temporal")]
    tiles_x = -(-n2 // cfg.OUT_COLS)
    tiles_y = -(-n1 // cfg.OUT_ROWS)
    slots = device_props.sm_count * cfg.blocks_per_sm

    def geometry(a0_begin, a0_end):
        n_units = a0_end - a0_begin
        chunks = plan_chunks(n_units, tiles_x * tiles_y, slots, 5, 32)
        forced = int(os.environ.get("STENCIL_CUDA_TEMPORAL_CHUNKS", "0"))        # tuning knob
        if forced:
            chunks = max(1, min(forced, n_units // 4))
        chunk_len = -(-n_units // chunks)
        chunks = -(-n_units // chunk_len)
        launch = Launch((tiles_x, tiles_y, chunks), (cfg.threads, 1, 1), cfg.smem_bytes)
        launch.extra = (chunk_len,)
        return launch

    plan = CudaPlan(program, domain, 'temporal', source, FUNCTION, options, geometry,
                    tensor_maps=specs, notes={'config': cfg.describe()})
    plan.function_no_peer = FUNCTION + "_interior" if domain.is_slab else None
    plan.function_step = FUNCTION + "_step" if (domain.is_slab and second is None) else None

    def step_geometry(a0_begin, a0_end):
        """grid and chunking of the one-launch step over the slab's real planes [a0_begin, a0_end):
        ``launch.ranges`` = (top chunk, middle range, bottom chunk); the chunk at each open end is a
        z-slice of its own, scheduled first (its first / last ghost-depth planes also go to the
        neighbour), the middle range is cut into chunks of ``launch.extra[0]`` planes"""
        launch = geometry(a0_begin, a0_end)
        edges = int(domain.open_lo) + int(domain.open_hi)
        planes = a0_end - a0_begin
        chunk_len, chunks = launch.extra[0], launch.grid[2]
        if chunks < edges:                        # a thin slab: one chunk per open end at least
            chunks = edges
            chunk_len = -(-planes // chunks)
            launch.extra = (chunk_len,)
        if planes < edges * domain.slab_ghost:
            raise StencilException("slab thinner than the planes it exchanges")
        top = (a0_begin, min(a0_begin + chunk_len, a0_end)) if domain.open_lo else (a0_begin, a0_begin)
        last_begin = a0_begin + (-(-planes // chunk_len) - 1) * chunk_len
        bottom = (max(last_begin, top[1]), a0_end) if domain.open_hi else (a0_end, a0_end)
        if domain.open_hi and bottom[1] - bottom[0] < domain.slab_ghost:
            # the last chunk is a short remainder: it must hold every plane the neighbour receives
            bottom = (a0_end - domain.slab_ghost, a0_end)
            top = (top[0], min(top[1], bottom[0]))
        middle = (top[1], bottom[0])
        middle_chunks = -(-(middle[1] - middle[0]) // chunk_len) if middle[1] > middle[0] else 0
        launch.grid = (launch.grid[0], launch.grid[1], max(1, edges + middle_chunks))
        launch.ranges = (top, middle, bottom)
        return launch
    plan.step_geometry = step_geometry
    return plan
