"""Strategy 'stream': TMA-staged shared-memory tiles streamed along axis 0.

The fast path for float32 grids of rank 2 and 3.  It supersedes the reference's OpenCL kernel
(``backend/ocl.py:363-594``: one work-item per point, a ``__local`` tile of the first argument
filled by a strided cooperative load with a div/mod and a per-dimension clamp per element, a
barrier, taps from local memory) and its per-dimension halo-copy kernels
(``backend/ocl_boundary_copier.py:97-134``).  The design, B200-first:

  * A CTA owns a tile of the fast axes (3-d: TY rows x TX columns of a plane; 2-d: TX columns)
    and marches along axis 0 over a chunk of planes (3-d) / rows (2-d).
  * One producer warp feeds a ring of S shared-memory stages with ``cp.async.bulk.tensor``
    (TMA) boxes that include the halo; completion is tracked by mbarriers, consumer warps hand
    stages back through a second set of mbarriers.  TMA zero-fills what lies outside the grid.
  * Each consumer thread computes a register block of BY rows x 4 consecutive columns per plane.
    Values needed from several planes of the sweep are loaded from shared memory ONCE, when the
    plane is newest, and kept in a rotating register window (the loop is emitted once per window
    phase so that rotation is pure renaming).  Stencils whose window would not fit in registers
    (bilateral filter, radius 9) read every tap row from the shared-memory ring instead.
  * All four boundary modes live in the same kernel.  Along axis 0 the producer clamps the TMA
    coordinate (rank 3) / the consumer clamps the row index (rank 2).  Inside the plane 'clamp'
    is applied while the consumers load their registers: halo rows are read through clamped row
    offsets and out-of-grid columns (zero-filled by TMA) are overwritten by a select cascade that
    only the warps of edge tiles execute.  'zero'/'copy' select 0 / the centre value for halo
    points at the store.  Every output point of the launch is written exactly once, with 128-bit
    stores.
  * Arithmetic is the typed per-point program, statement for statement, with FMA contraction
    off -- bit-identical to the direct strategy and to the reference's C.

Algorithmic traffic: 8 B per point (one float32 read, one write).  Roofline: HBM bandwidth.
"""
from ...stencil_exception import StencilException
from .. import point_program as pp
import re

from .codegen_common import PROLOGUE, emit_statements

_LOCAL_NAME = re.compile(r'\b_t_[A-Za-z0-9_]+\b')

FUNCTION = "stencil_stream"

PTX_HELPERS = r'''
struct alignas(64) TensorMap { u64 opaque[16]; };

__device__ __forceinline__ u32 sc_smem(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sc_smem(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sc_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(sc_smem(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(sc_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const TensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(tm) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap* tm, u64* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(sc_smem(dst)), "l"(tm), "r"(sc_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* tm, u64* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(sc_smem(dst)), "l"(tm), "r"(sc_smem(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float lds_f32(u32 address) {       // read-only tables: not volatile, may be scheduled freely
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(address));
    return v;
}
__device__ __forceinline__ void st_global_v4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global@STOP@.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
'''


class TensorMapSpec(object):
    """What the runtime needs to encode one TMA descriptor (``sc_tensormap_tiled``)."""
    def __init__(self, arg, dims, strides_bytes, box, dtype_code=0, tag=""):
        self.arg = arg
        self.dims = tuple(int(d) for d in dims)
        self.strides_bytes = tuple(int(s) for s in strides_bytes)
        self.box = tuple(int(b) for b in box)
        self.dtype_code = dtype_code
        self.key = (arg, self.dims, self.strides_bytes, self.box, dtype_code, tag)


class StreamConfig(object):
    """Tile / pipeline shape of one generated streaming kernel."""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def describe(self):
        keys = ("rank", "TX", "TY", "BY", "U", "S", "NCW", "PADX", "HY", "HZ_LO", "HZ_HI", "W",
                "mode", "blocks_per_sm", "smem_bytes")
        return {k: getattr(self, k) for k in keys}


# ---- applicability ------------------------------------------------------------------------------
def _grid_args(program):
    return [a for a in program.args if a.used_as_grid]


def applicable(program, domain, device_props, forced=False):
    if program.dim not in (2, 3):
        return False, "rank {} (stream handles 2 and 3)".format(program.dim)
    if program.out_ctype != pp.FLOAT:
        return False, "output is not float32"
    for arg in _grid_args(program):
        if arg.ctype != pp.FLOAT:
            return False, "grid argument '{}' is not float32".format(arg.name)
    n_fast = program.shape[-1]
    if program.boundary == 'periodic':
        return False, "periodic boundaries (TMA tiles cannot wrap around the grid)"
    if domain.row_pitch % 4 != 0:
        return False, "rows of {} floats are not a multiple of 16 bytes (TMA stride rule)".format(domain.row_pitch)
    if n_fast < 32:
        return False, "rows shorter than 32 floats"
    if program.dim == 3 and program.shape[1] < 4:
        return False, "fewer than 4 rows per plane"
    if program.shape[0] < 2:
        return False, "a single plane/row along the streaming axis"
    if program.npoints < 4096:
        return False, "tiny grid"
    if program.npoints < (1 << 20) and not forced:
        # measured on B200 (us per dependent sweep, floor 4.1): 512^2 and 64^3 take 4.1 with one
        # thread per point and 4.1-6.2 through the TMA pipeline; from 1024^2 (5.1 vs 6.2) and
        # 128^3 (6.2 vs 10.3) on the pipeline wins
        return False, "small grid ({} points): launch-latency bound, direct is faster".format(program.npoints)
    lo, hi = program.tap_extent()
    if max(-lo[-1], hi[-1]) > 12:
        return False, "tap radius along the fast axis above 12"
    if any(s >= 2 ** 31 for s in program.shape):
        return False, "dimension above 2^31"
    try:
        if wants_plane_embedding(program, domain):
            from .domain import Domain
            lifted = embed_as_plane(program)
            choose_config(lifted, Domain(lifted.shape, lifted.ghost_depth), device_props, {'embedded': True})
        else:
            choose_config(program, domain, device_props)
    except StencilException as error:
        return False, str(error)
    return True, "float32 rank-{} grid with 16-byte rows".format(program.dim)


# ---- rank-2 stencils with a large window run as ONE plane of the rank-3 path (tiled, no streaming)
def embed_as_plane(program):
    """A view of a rank-2 program as rank 3 with a single plane: shape (1, n0, n1), taps (0, dy, dx).
    Used when the row window is too large for registers (bilateral filter): such stencils are
    compute bound, so a 2-d tile with a full halo and many resident CTAs beats row streaming."""
    import copy as _copy

    def lift(expr):
        if isinstance(expr, pp.Tap):
            return pp.Tap(expr.arg, (0,) + tuple(expr.offset), expr.ctype)
        clone = _copy.copy(expr)
        for name in ('left', 'right', 'operand', 'index'):
            if hasattr(clone, name):
                setattr(clone, name, lift(getattr(clone, name)))
        if hasattr(clone, 'args') and isinstance(clone, pp.Call):
            clone.args = [lift(a) for a in clone.args]
        return clone

    statements = []
    for statement in program.statements:
        clone = _copy.copy(statement)
        clone.expr = lift(statement.expr)
        statements.append(clone)
    lifted = pp.PointProgram(3, (1,) + tuple(program.shape), program.out_ctype, program.args, statements,
                             (0,) + tuple(program.ghost_depth), program.boundary)
    return lifted


def wants_plane_embedding(program, domain):
    if program.dim != 2 or domain.is_slab:
        return False
    return Footprint(program, 1, program.boundary == 'copy').window_registers > 160


# ---- footprint analysis -------------------------------------------------------------------------
def _tap3(offset):
    """(d_stream, d_row, d_col) of a tap offset for rank 2 or 3."""
    if len(offset) == 3:
        return offset
    return (offset[0], 0, offset[1])


class Footprint(object):
    """Which shared-memory values one thread needs, and at which ages of the streamed plane.

    ``age`` is the axis-0 offset of a plane relative to the output plane (the tap's first
    component).  A value is identified by (grid, row, col) relative to the thread's register
    block; ``ages[(g, row, col)]`` is the set of ages at which it is used.
    """
    def __init__(self, program, BY, copy_center):
        self.ages = {}
        for tap in program.taps():
            dz, dy, dx = _tap3(tap.offset)
            for b in range(BY):
                for j in range(4):
                    self.ages.setdefault((tap.arg, b + dy, j + dx), set()).add(dz)
        if copy_center:
            for b in range(BY):
                for j in range(4):
                    self.ages.setdefault((0, b, j), set()).add(0)
        # rows of the grid that end inside a thread's quad of columns (length no multiple of 4:
        # only reachable through the padded-row buffers of iterate())
        self.ragged = program.shape[-1] % 4 != 0
        if program.boundary == 'clamp':
            # a column left of the grid takes the value of its right-hand neighbour (and vice versa
            # on the right), cascading up to the thread's own columns: those must be loaded too
            # (ragged rows: the cascade on the right can start inside the quad, at column 0)
            first_right = 0 if self.ragged else 3
            for (g, row, col), ages in list(self.ages.items()):
                inner = range(col + 1, 1) if col < 0 else range(first_right, col) if col > 3 else ()
                for c in inner:
                    self.ages.setdefault((g, row, c), set()).update(ages)
        all_ages = [a for s in self.ages.values() for a in s] or [0]
        self.hz_lo, self.hz_hi = min(min(all_ages), 0), max(max(all_ages), 0)
        self.window = self.hz_hi - self.hz_lo + 1
        self.window_registers = sum(max(s) - min(s) + 1 for s in self.ages.values())
        # age at which the last first-load of a plane happens: the plane may leave smem after it
        self.release_age_regwin = min(max(s) for s in self.ages.values())


def choose_config(program, domain, device_props, override=None):
    """Pick tile shape, register block, ring depth and occupancy for a stencil and device."""
    rank = program.dim
    lo, hi = program.tap_extent()
    hx = max(-lo[-1], hi[-1])
    hy = max(-lo[1], hi[1]) if rank == 3 else 0
    PADX = 4 * ((hx + 3) // 4) if hx else 0
    copy_center = program.boundary == 'copy'
    n_fast = program.shape[-1]
    n_grids = len(_grid_args(program))
    override = dict(override or {})
    # tuning knob: STENCIL_CUDA_STREAM="TY=64,BY=4,NCW=16,prefetch=2" forces tile parameters
    import os
    for item in filter(None, os.environ.get("STENCIL_CUDA_STREAM", "").split(",")):
        key, _, value = item.partition("=")
        override[key.strip()] = value.strip() if key.strip() == 'mode' else int(value)

    candidates = []
    embedded = bool(override.pop('embedded', False))
    if embedded:
        for TY, BY, NCW in ((16, 2, 8), (32, 4, 8), (32, 2, 16), (16, 4, 4), (8, 1, 8)):
            candidates.append(dict(TX=128, TY=TY, BY=BY, NCW=NCW, U=1, prefetch=0))
    elif rank == 3:
        # measured on B200 (GStencil/s; Laplacian 512^3 / Jacobi 2048x1024^2 / 256x1024^2 slab):
        # (64,4,16,p2) 765 / 775 / 784, (64,4,16,p3) 729 / 766 / 768, (32,4,8,p3) 731 / 731 / 770,
        # (32,4,8,p2) 700 / 671 / 681, (16,4,4,p3) 719 / 708 / 759.  Ties in the traffic model go
        # to the earlier candidate.
        for TY, BY, NCW, prefetch in ((64, 4, 16, 2), (64, 4, 16, 3), (32, 4, 8, 3), (32, 4, 8, 2), (16, 4, 4, 3),
                                      (16, 2, 8, 3), (8, 2, 4, 3), (8, 1, 8, 3)):
            candidates.append(dict(TX=128, TY=TY, BY=BY, NCW=NCW, U=1, prefetch=prefetch))
    else:
        # rank 2: every consumer warp owns a 128-column sub-tile with its own TMA box and halo
        # BY > 1: a thread owns BY four-column blocks, one in each of BY sub-tiles of its warp
        # (halves the per-point cost of the row/stage bookkeeping, which bounds 5x5-like stencils)
        # Listed in measured order (B200, conv 5x5 16384^2 clamp, GStencil/s): (4,4,2) 764,
        # (4,4,1) 678, (4,8,1) 672, (2,8,2) 538, (4,8,2) 509; 3x3 8192^2: (4,4,2) 680, (4,8,1) 665.
        # Among the candidates that waste the fewest lanes on the grid's width the first one wins.
        for NCW, U, BY in ((4, 4, 2), (4, 8, 1), (4, 4, 1), (2, 8, 1), (2, 4, 1), (8, 8, 1), (4, 16, 1),
                           (2, 16, 1), (1, 16, 1), (1, 8, 1), (1, 4, 1)):
            candidates.append(dict(TX=NCW * BY * 128, TY=1, BY=BY, NCW=NCW, U=U, prefetch=2))
    best = None
    if any(k in override for k in ('TY', 'BY', 'NCW', 'U')):
        forced = dict(candidates[0])
        forced.update({k: v for k, v in override.items() if k in forced})
        if rank == 2:
            forced['TX'] = forced['NCW'] * forced['BY'] * 128
        candidates = [forced]
    elif 'prefetch' in override:
        for cand in candidates:
            cand['prefetch'] = override['prefetch']
    n_stream = domain.sweep_range[1] - domain.sweep_range[0]
    for cand in candidates:
        fp = Footprint(program, cand['BY'], copy_center)
        mode = override.get('mode') or ('regwin' if fp.window_registers <= 160 else 'smem')
        if embedded:
            mode = 'smem'
        W = fp.window
        if mode == 'regwin':
            resident_units = fp.hz_hi - fp.release_age_regwin + 1
        else:
            resident_units = W
        U = cand['U']
        if rank == 2 and U <= max(-fp.hz_lo, fp.hz_hi):
            continue
        resident_stages = (resident_units + U - 1) // U + (1 if U > 1 else 0)
        S = resident_stages + cand['prefetch']
        pitch = 128 + 2 * PADX
        if pitch > 256:
            continue
        rows = (cand['TY'] + 2 * hy) if rank == 3 else U
        n_sub = 1 if rank == 3 else cand['NCW'] * cand['BY']
        stage_floats = n_sub * rows * pitch
        stage_bytes = stage_floats * 4
        stage_bytes_aligned = (stage_bytes + 127) // 128 * 128
        smem = n_grids * S * stage_bytes_aligned + 3 * S * 8 + 128
        static_smem = sum(a.size * 128 for a in table_plan(program, embedded)[3].values())
        smem_total = smem + static_smem
        threads = (cand['NCW'] + 1) * 32
        if smem_total > device_props.smem_per_block_optin:
            continue
        blocks_per_sm = min(device_props.smem_per_sm // (smem_total + 1024),
                            device_props.max_threads_per_sm // threads, 8)
        if mode == 'smem':
            # tap rows are re-read from smem per statement group: about one row of the footprint
            # per register-block row stays live, plus accumulators and hoisted table entries
            row_values = len({k[2] for k in fp.ages})
            registers = 40 + cand['BY'] * (row_values + 4) + len(table_plan(program, embedded)[1])
            blocks_per_sm = min(blocks_per_sm, max(1, device_props.regs_per_sm // (threads * registers)))
        if 'blocks' in override:
            blocks_per_sm = min(blocks_per_sm, override['blocks'])
        if blocks_per_sm < 1:
            continue
        # cost model: DRAM traffic per useful byte (halo rows, 32-byte sectors of the column pad,
        # window start-up per chunk; L2 does not absorb the overlap -- measured, profiles/) divided
        # by how well tiles x chunks fill whole waves of resident CTAs, with lanes wasted on
        # partial tiles counted as extra work
        tiles_x = -(-n_fast // cand['TX'])
        waste = tiles_x * cand['TX'] / float(n_fast)
        tiles = tiles_x
        read_factor = (128 * 4 + (64 if PADX else 0)) / 512.0
        if rank == 3:
            n_rows = program.shape[1]
            tiles_y = -(-n_rows // cand['TY'])
            tiles *= tiles_y
            waste *= tiles_y * cand['TY'] / float(n_rows)
            read_factor *= (cand['TY'] + 2 * hy) / float(cand['TY'])
        slots = device_props.sm_count * blocks_per_sm
        min_chunk = chunk_floors(rank, W, U, override)
        chunks = plan_chunks_adaptive(n_stream, tiles, slots, W, min_chunk, device_props.sm_count)
        chunk_len = -(-n_stream // chunks)
        read_factor *= (chunk_len + W - 1) / float(chunk_len)
        items = tiles * (-(-n_stream // chunk_len))
        if items >= slots:
            fill = items / float(-(-items // slots) * slots)
        else:
            # a grid too small to fill the resident slots: what counts is that every SM has work
            sms = device_props.sm_count
            fill = items / float(-(-items // sms) * sms)
        resident_warps = blocks_per_sm * cand['NCW']
        score = waste * (1.0 + read_factor) / 2.0 / (fill ** 0.25) * (1.0 if resident_warps >= 12 else 1.1)
        if fill < 0.8:
            score *= (0.8 / fill) ** 0.75
        if rank == 3 and chunk_len < 32:
            # short chunks: the pipeline start-up (about two plane times) shows unless a second
            # resident CTA covers it.  Measured at 256^3: (64,4,16) one CTA/SM 680 GStencil/s,
            # (32,4,8) two CTAs/SM 746-764
            score *= 1.0 + 2.0 / chunk_len / min(blocks_per_sm, 2)
        if rank == 2 and not any(k in override for k in ('TY', 'BY', 'NCW', 'U')):
            few_warps = {1: 1.3, 2: 1.1}.get(cand['NCW'], 1.0)
            # lane waste, then measured order; a candidate that leaves SMs without a CTA loses to
            # one that does not (1024^2 Jacobi: (4,4,2) 128 CTAs 6.2 us, (4,4,1) 256 CTAs 5.1 us)
            starved = 1.1 if items < device_props.sm_count else 1.0
            score = waste * few_warps * starved * 100 + 0.5 * candidates.index(cand)
        if embedded:
            score = float(len(candidates) - candidates.index(cand)) * -1.0      # first that fits wins
        config = StreamConfig(rank=rank, PADX=PADX, HX=hx, HY=hy, HZ_LO=fp.hz_lo, HZ_HI=fp.hz_hi, W=W,
                              embedded=embedded,
                              mode=mode, S=S, PITCH=pitch, ROWS=rows, STAGE_FLOATS=stage_bytes_aligned // 4,
                              STAGE_BYTES=stage_bytes, smem_bytes=smem, threads=threads, NSUB=n_sub,
                              SUB_FLOATS=rows * pitch,
                              blocks_per_sm=blocks_per_sm, footprint=fp, n_grids=n_grids,
                              score=score, TX=cand['TX'], TY=cand['TY'], BY=cand['BY'], NCW=cand['NCW'],
                              U=cand['U'], min_chunk=min_chunk)
        if best is None or score < best.score - 1e-9:
            best = config
    if best is None:
        raise StencilException("no streaming tile fits this stencil in shared memory")
    return best


def chunk_floors(rank, window, units_per_stage, override=None):
    """(usual, small-grid) lower bounds of the axis-0 chunk length.  Long chunks keep the re-read of
    the window-1 start-up planes (rows) per chunk small, which matters when they come from DRAM;
    a grid that cannot give every SM a CTA at that length is L2-resident, and there short chunks
    win: measured on B200, Jacobi 128^3 16.8 -> 6.2 us/sweep, 1024^2 14.4 -> 5.1 us/sweep."""
    forced = (override or {}).get('min_chunk')
    if forced:
        return forced, forced
    if rank == 3:
        return max(4 * window, 16), window + 1
    return max(4 * units_per_stage, 64), 2 * units_per_stage


def plan_chunks_adaptive(n_units, tiles, slots, window, floors, sm_count):
    usual, small = floors
    chunks = plan_chunks(n_units, tiles, slots, window, usual)
    if tiles * chunks < sm_count and small < usual:
        chunks = plan_chunks(n_units, tiles, slots, window, small)
    return chunks


def plan_chunks(n_units, tiles, slots, window, min_chunk):
    """Number of axis-0 chunks so that tiles*chunks fills the resident CTA slots in whole waves
    while the window start-up (window-1 extra units per chunk) stays small."""
    best_chunks, best_score = 1, -1.0
    max_chunks = max(1, n_units // max(min_chunk, 1))
    for chunks in range(1, max_chunks + 1):
        length = -(-n_units // chunks)
        if chunks > 1 and length < min_chunk:
            break
        items = tiles * (-(-n_units // length))
        waves = -(-items // slots)
        fill = items / float(waves * slots)
        overhead = length / float(length + window - 1)
        score = fill * overhead
        if score > best_score + 1e-9:
            best_chunks, best_score = chunks, score
    return best_chunks


# ---- code generation ----------------------------------------------------------------------------
class _Emitter(object):
    def __init__(self):
        self.lines = []
        self.level = 0

    def __call__(self, text=""):
        self.lines.append(("    " * self.level + text) if text else "")

    def open(self, text):
        self(text + " {")
        self.level += 1

    def close(self, suffix=""):
        self.level -= 1
        self("}" + suffix)

    def text(self):
        return "\n".join(self.lines) + "\n"


def _group_loads(cols):
    """Split a set of needed columns (relative to the thread's first column) into aligned
    128-bit / 64-bit / 32-bit shared-memory loads.  Returns [(first_col, width, used_cols)]."""
    loads = []
    quads = sorted({c // 4 for c in cols})
    for q in quads:
        used = sorted(c for c in cols if c // 4 == q)
        if len(used) >= 3:
            loads.append((4 * q, 4, used))
        elif len(used) == 2 and used[0] % 2 == 0 and used[1] == used[0] + 1:
            loads.append((used[0], 2, used))
        else:
            for c in used:
                loads.append((c, 1, [c]))
    return loads


def generate(program, domain, cfg):
    import os
    rank = cfg.rank
    fp = cfg.footprint
    grids = _grid_args(program)
    ring_of = {arg.index: k for k, arg in enumerate(grids)}        # ring index per grid argument
    shape = program.shape
    N0 = shape[0]
    N1 = shape[1] if rank == 3 else 1
    N2 = shape[-1]
    TX, TY, BY, U, S, NCW = cfg.TX, cfg.TY, cfg.BY, cfg.U, cfg.S, cfg.NCW
    PADX, HX, HY, PITCH, ROWS = cfg.PADX, cfg.HX, cfg.HY, cfg.PITCH, cfg.ROWS
    HZ_LO, HZ_HI, W = cfg.HZ_LO, cfg.HZ_HI, cfg.W
    SF = cfg.STAGE_FLOATS
    NSUB, SUBF = cfg.NSUB, cfg.SUB_FLOATS
    regwin = cfg.mode == 'regwin'
    phases = W if regwin else 1
    boundary = program.boundary
    clamp = boundary == 'clamp'
    copy = boundary == 'copy'
    c_lo, c_hi = domain.clamp_lo, domain.clamp_hi
    i_lo, i_hi = domain.interior_lo, domain.interior_hi
    col_axis = rank - 1
    release_age = fp.release_age_regwin if regwin else HZ_LO
    rel_back = HZ_HI - release_age

    # tables: constant-index entries are hoisted into registers, data-indexed small tables go to
    # smem -- bank-replicated (entry k of lane l at word k*32+l: conflict-free gathers) in the
    # compute-bound tiled kernels
    embedded = bool(getattr(cfg, 'embedded', False))
    data_tables, const_refs, smem_tables, replicated = table_plan(program, embedded)
    guard_halo_body = (not clamp) and bool(data_tables)

    def tag(v):
        return "m{}".format(-v) if isinstance(v, int) and v < 0 else str(v)

    def value_name(g, where, row, col):
        return "v{}_{}_{}_{}".format(g, tag(where), tag(row), tag(col))

    def slot_of(phase, age, key):
        ages = fp.ages[key]
        if max(ages) == min(ages):
            return 0
        return (phase + age - HZ_HI) % W

    def name_of(phase, g, age, row, col):
        if regwin:
            return value_name(g, slot_of(phase, age, (g, row, col)), row, col)
        return value_name(g, "a" + tag(age), row, col)

    # rank 2: a thread's block b lives in sub-tile (warp*BY + b): "row" stride = one sub-tile,
    # and its columns start 128*b to the right of block 0
    row_stride = PITCH if rank == 3 else SUBF

    def col_base(b):
        return "gx" if rank == 3 or b == 0 else "gx + {}".format(128 * b)

    roll = find_rolled_run(program) if (not regwin and rank == 3) else None
    if roll is not None and os.environ.get("STENCIL_CUDA_NO_ROLL"):
        roll = None

    e = _Emitter()
    e.lines.append(PROLOGUE)
    e.lines.append(PTX_HELPERS.replace("@STOP@", os.environ.get("STENCIL_CUDA_STORE_OP", ".cs")))
    if roll is not None:
        e("// per-tap parameters of the re-rolled neighbor loop ({} rows x {} taps)".format(
            len(roll['dys']), len(roll['dxs'])))
        for slot, param in enumerate(roll['params']):
            if param is None:
                continue
            if param[0] == 'const':
                values = ", ".join(pp.Const(v, param[1]).c_text() for v in param[2])
                e("__constant__ {} cpar{}[{}] = {{{}}};".format(param[1], slot, len(param[2]), values))
            else:
                e("__constant__ int ipar{}[{}] = {{{}}};".format(slot, len(param[3]), ", ".join(map(str, param[3]))))
    e("// stream strategy: rank {} grid {} boundary {} | tile TX={} TY={} BY={} U={} ring S={} "
      "window W={} ({})".format(rank, shape, boundary, TX, TY, BY, U, S, W, cfg.mode))
    params = ["const {}* __restrict__ a{}".format(arg.ctype, arg.index) for arg in program.args]
    for arg in grids:
        params.append("const __grid_constant__ TensorMap tm{}".format(arg.index))
    params += ["float* __restrict__ out", "float* __restrict__ out_peer", "i64 peer_shift",
               "int a0_begin", "int a0_end", "int chunk_len"]
    e('extern "C" __global__ void __launch_bounds__({}, {})'.format(cfg.threads, cfg.blocks_per_sm))
    e("{}({})".format(FUNCTION, ", ".join(params)))
    e.open("")
    e("extern __shared__ __align__(128) unsigned char smem_raw[];")
    e("float* const ring = reinterpret_cast<float*>(smem_raw);")
    e("u64* const bar_full = reinterpret_cast<u64*>(smem_raw + {});".format(cfg.n_grids * S * SF * 4))
    e("u64* const bar_empty = bar_full + {};".format(S))
    for index, arg in sorted(smem_tables.items()):
        e("__shared__ {} tab{}[{}];".format(arg.ctype, index, arg.size))
    for index, arg in sorted(replicated.items()):
        e("__shared__ float rep{}[{}];".format(index, arg.size * 32))
    if roll is not None:
        for slot, param in enumerate(roll['params']):
            if param is not None and param[0] == 'table':
                e("__shared__ {} spar{}[{}];".format(param[1], slot, len(param[3])))
    e("const int tid = threadIdx.x;")
    e("const int warp = tid >> 5;")
    e("const int lane = tid & 31;")
    e("const int x0 = blockIdx.x * {};".format(TX))
    if embedded:
        e("// one plane: a0_begin/a0_end select the ROWS this launch writes (tiles start at a0_begin)")
        e("const int y0 = a0_begin + blockIdx.y * {};".format(TY))
        e("const int row_end = a0_end;")
        e("const int u_begin = 0;")
        e("const int u_end = 1 + 0 * chunk_len;")
    else:
        if rank == 3:
            e("const int y0 = blockIdx.y * {};".format(TY))
            e("const int row_end = {};".format(N1))
        e("const int u_begin = a0_begin + blockIdx.z * chunk_len;")
        e("const int u_end = min(u_begin + chunk_len, a0_end);")
    e("if (u_begin >= u_end) return;")
    e("const int p_first = u_begin + ({});          // axis-0 index of streamed unit 0".format(HZ_LO))
    e("const int n_units = (u_end - u_begin) + {};".format(W - 1))
    e("const int n_stages = (n_units + {}) / {};".format(U - 1, U))
    e.open("if (tid == 0)")
    e.open("for (int s = 0; s < {}; ++s)".format(S))
    e("mbar_init(&bar_full[s], 1);")
    e("mbar_init(&bar_empty[s], {});".format(NCW))
    e.close()
    e("fence_barrier_init();")
    e.close()
    for index, arg in sorted(smem_tables.items()):
        e("for (int k = tid; k < {n}; k += {t}) tab{i}[k] = a{i}[k];".format(n=arg.size, t=cfg.threads, i=index))
    for index, arg in sorted(replicated.items()):
        e("for (int k = tid; k < {n}; k += {t}) rep{i}[k] = a{i}[k >> 5];".format(
            n=arg.size * 32, t=cfg.threads, i=index))
    if roll is not None:
        for slot, param in enumerate(roll['params']):
            if param is not None and param[0] == 'table':
                e("for (int k = tid; k < {n}; k += {t}) spar{s}[k] = a{a}[ipar{s}[k]];".format(
                    n=len(param[3]), t=cfg.threads, s=slot, a=param[2]))
    for index, arg in sorted(replicated.items()):
        e("const u32 repk{i} = sc_smem(rep{i}) + (u32)lane * 4u - 0x80000000u;".format(i=index))
    e("__syncthreads();")
    e("pdl_wait();                         // the grids may still be written by the preceding sweep")
    e()

    # ================================================================== producer warp
    e.open("if (warp == {})".format(NCW))
    for arg in grids:
        e("if (lane == 0) tma_prefetch_desc(&tm{});".format(arg.index))
    e.open("for (int k = 0; k < n_stages; ++k)")
    e("const int s = k % {};".format(S))
    e("if (k >= {S}) mbar_wait(&bar_empty[s], ((k / {S}) - 1) & 1);".format(S=S))
    e.open("if (lane == 0)")
    e("u64* const bar = &bar_full[s];")
    e("mbar_expect_tx(bar, {});".format(cfg.STAGE_BYTES * cfg.n_grids))
    if rank == 2:
        e("const int r0 = p_first + k * {};   // rows outside the grid are zero-filled and never read".format(U))
    for arg in grids:
        dst = "ring + ({} + s) * {}".format(ring_of[arg.index] * S, SF)
        if rank == 3:
            e("tma_load_3d({dst}, &tm{g}, bar, x0 - {px}, y0 - {hy}, SC_CLAMP(p_first + k, {lo}, {hi}));".format(
                dst=dst, g=arg.index, px=PADX, hy=HY, lo=c_lo[0], hi=c_hi[0]))
        else:
            e.open("for (int w = 0; w < {}; ++w)".format(NSUB))
            e("float* const dst = {} + w * {};".format(dst, SUBF))
            e("const int xs = x0 + w * 128 - {};".format(PADX))
            e("tma_load_2d(dst, &tm{g}, bar, xs, r0);".format(g=arg.index))
            e.close()
    e.close()
    e.close()   # for k
    e("// all of this CTA's loads have landed: let the next sweep's CTAs be scheduled.  (Earlier would")
    e("// place them while this grid still fills the SMs -- measured: unevenly, 20-35 % slower.)")
    e.lines.append("#ifdef SC_PDL")
    e("mbar_wait(&bar_full[(n_stages - 1) % {S}], ((n_stages - 1) / {S}) & 1);".format(S=S))
    e("pdl_launch_dependents();")
    e.lines.append("#endif")
    e("return;")
    e.close()   # producer
    e()

    # ================================================================== consumer warps
    if rank == 3:
        e("const int tq = lane;                 // quad (4 columns) within the tile row")
        e("const int tr = warp;                 // group of {} rows".format(BY))
        e("const int gy = y0 + tr * {};".format(BY))
        e("const int sbase = ({} + tr * {}) * {} + {} + tq * 4;".format(HY, BY, PITCH, PADX))
    else:
        e("const int tq = warp * {} + lane;      // quad of block 0 within the CTA's tile".format(32 * BY))
        e("const int sbase = warp * {} + {} + lane * 4;   // block 0's 128-column sub-tile".format(BY * SUBF, PADX))
    e("const int gx = x0 + tq * 4;          // first of this thread's 4 columns (block 0)")
    if rank == 3:
        e("const bool col_ok = gx < {};".format(N2))
    else:
        for b in range(BY):
            e("const bool col_ok{} = {} < {};".format(b, col_base(b), N2))
    clamp_rows = clamp and rank == 3 and (HY > 0 or BY > 1)
    clamp_cols = clamp and HX > 0
    if clamp_rows:
        e("// 'clamp' inside the plane is applied while loading: rows through clamped row offsets,")
        e("// columns by a select cascade that only warps of edge tiles execute")
        for row in sorted({k[1] for k in fp.ages}):
            e("const int ro{t} = (min(max(gy + ({r}), 0), {n}) - (gy + ({r}))) * {p};".format(
                t=tag(row), r=row, n=N1 - 1, p=PITCH))
    if clamp_cols:
        if rank == 3:
            e("const bool x_edge = x0 == 0 || x0 + {} > {};".format(TX + HX, N2))
        else:
            e("const bool x_edge = x0 + warp * {w} == 0 || x0 + warp * {w} + {r} > {n};".format(
                w=128 * BY, r=128 * BY + HX, n=N2))
    if not clamp:
        e("u32 halo_mask = 0;                   // bit (b*4+j): the point lies in the in-plane halo")
        for b in range(BY):
            for j in range(4):
                tests = []
                if rank == 3:
                    if i_lo[1] > 0:
                        tests.append("gy + {} < {}".format(b, i_lo[1]))
                    if i_hi[1] < N1:
                        tests.append("gy + {} >= {}".format(b, i_hi[1]))
                if i_lo[col_axis] > 0:
                    tests.append("{} + {} < {}".format(col_base(b), j, i_lo[col_axis]))
                if i_hi[col_axis] < N2:
                    tests.append("{} + {} >= {}".format(col_base(b), j, i_hi[col_axis]))
                if tests:
                    e("if ({}) halo_mask |= {}u;".format(" || ".join(tests), 1 << (b * 4 + j)))
    for (arg, index), ctype in sorted(const_refs.items()):
        e("const {} tc{}_{} = a{}[{}];".format(ctype, arg, index, arg, index))

    if regwin:
        names = []
        for (g, row, col), ages in sorted(fp.ages.items()):
            slots = W if max(ages) > min(ages) else 1
            names += [value_name(g, slot, row, col) for slot in range(slots)]
        for first in range(0, len(names), 8):
            e("float {};".format(", ".join(n + " = 0.0f" for n in names[first:first + 8])))

    counter = [0]

    def emit_column_fix(names_by_col, row=0):
        """cascade the nearest in-grid column outward (names_by_col: col -> register name)"""
        base = col_base(row if rank == 2 else 0)
        left = sorted((c for c in names_by_col if c < 0), reverse=True)
        right = sorted(c for c in names_by_col if c > (0 if fp.ragged else 3))
        for c in left:
            e("if ({} + ({}) < 0) {} = {};".format(base, c, names_by_col[c], names_by_col[c + 1]))
        for c in right:
            e("if ({} + {} > {}) {} = {};".format(base, c, N2 - 1, names_by_col[c], names_by_col[c - 1]))

    def emit_loads(phase, g, age, keys):
        """shared memory -> registers for the (row, col) values `keys` of grid g at `age`"""
        pointer = "u{}_{}".format(g, tag(age))
        decl = "" if regwin else ("float " if clamp_cols else "const float ")
        by_row = {}
        for (_, row, col) in keys:
            by_row.setdefault(row, set()).add(col)
        for row in sorted(by_row):
            for first, width, used in _group_loads(by_row[row]):
                offset = "{}".format(row * row_stride + first)
                if clamp_rows:
                    offset = "ro{} + ({})".format(tag(row), offset)
                if width == 1:
                    e("{}{} = {}[{}];".format(decl, name_of(phase, g, age, row, used[0]), pointer, offset))
                    continue
                counter[0] += 1
                tmp = "q{}".format(counter[0])
                vec = "float4" if width == 4 else "float2"
                e("const {v} {t} = *reinterpret_cast<const {v}*>({p} + ({o}));".format(v=vec, t=tmp, p=pointer, o=offset))
                e(" ".join("{}{} = {}.{};".format(decl, name_of(phase, g, age, row, c), tmp, "xyzw"[c - first])
                           for c in used))
        n_outer = sum(1 for cols in by_row.values() for c in cols if c < 0 or c > 3)
        if clamp_cols and n_outer:
            # few selects: leave them unconditional (loop-invariant predicates, no basic-block split
            # between the loads and the arithmetic); many: only edge-tile warps execute them
            branch = n_outer > int(os.environ.get("STENCIL_CUDA_FIX_BRANCH", "8"))
            e.open("if (x_edge)" if branch else "")
            for row in sorted(by_row):
                outer = [c for c in by_row[row] if c < 0 or c > 3]
                if not outer:
                    continue
                span = range(min(min(outer), 0), max(max(outer), 3) + 1)
                emit_column_fix({c: name_of(phase, g, age, row, c) for c in span if (g, row, c) in fp.ages}, row)
            e.close()

    current_tap_text, current_acc = [None], [None]

    def emit_rolled(phase, loaded, acc):
        """the re-rolled neighbor loop: runtime loop over tap rows, taps of a row unrolled"""
        import copy as _copy
        template = roll['template']
        taps, pures = [], []
        _shape_of(template.expr, taps, pures)
        moving_ids = {id(taps[k]) for k in roll['moving']}
        slot_counter = [0]

        def substitute(expr):
            if isinstance(expr, pp.Tap):
                return expr
            if _is_pure(expr):
                slot = slot_counter[0]
                slot_counter[0] += 1
                if roll['params'][slot] is None:
                    return expr
                return pp.LocalRef("@P{}@".format(slot), expr.ctype)
            clone = _copy.copy(expr)
            for name in ('left', 'right', 'operand', 'index'):
                if hasattr(clone, name):
                    setattr(clone, name, substitute(getattr(clone, name)))
            if isinstance(clone, pp.Call):
                clone.args = [substitute(a) for a in clone.args]
            return clone
        # children() order must match _shape_of: Bin(left,right), Neg/Cast(operand), Call(args), TableRef(index)
        body = pp.Store(template.op, substitute(template.expr))
        dz, dys, dxs, gi = roll['dz'], roll['dys'], roll['dxs'], roll['grid']
        n_dx = len(dxs)
        # taps that do not move with the neighbor (e.g. the centre value) are loaded once
        wanted = {}
        for tap in taps:
            if id(tap) not in moving_ids:
                tz, ty, tx = _tap3(tap.offset)
                wanted.setdefault((tap.arg, tz), set()).update({b + ty for b in range(BY)})
        for (ga, age), rows in sorted(wanted.items()):
            keys = [k for k, ages in fp.ages.items()
                    if k[0] == ga and age in ages and k[1] in rows and (ga, age, k[1], k[2]) not in loaded]
            if keys:
                emit_loads(phase, ga, age, keys)
                loaded.update((ga, age, k[1], k[2]) for k in keys)
        e("#pragma unroll 1")
        e.open("for (int tr_ = 0; tr_ < {}; ++tr_)".format(len(dys)))
        e("const int pk = tr_ * {};".format(n_dx))
        cols = sorted({j + dx for j in range(4) for dx in dxs})
        if clamp_cols:
            cols = list(range(min(cols[0], 0), max(cols[-1], 3) + 1))
        decl = "float" if clamp_cols else "const float"
        for b in range(BY):
            if clamp_rows:
                e("const float* const rp{b} = u{g}_{z} + (min(max(gy + ({r}) + tr_, 0), {n}) - gy) * {p};".format(
                    b=b, g=gi, z=tag(dz), r=b + dys[0], n=N1 - 1, p=PITCH))
            else:
                e("const float* const rp{b} = u{g}_{z} + ({r} + tr_) * {p};".format(
                    b=b, g=gi, z=tag(dz), r=b + dys[0], p=PITCH))
            for first, width, used in _group_loads(set(cols)):
                if width == 1:
                    e("{} rw{}_{} = rp{}[{}];".format(decl, b, tag(used[0]), b, first))
                    continue
                counter[0] += 1
                tmp = "q{}".format(counter[0])
                vec = "float4" if width == 4 else "float2"
                e("const {v} {t} = *reinterpret_cast<const {v}*>(rp{b} + ({o}));".format(v=vec, t=tmp, b=b, o=first))
                e(" ".join("{} rw{}_{} = {}.{};".format(decl, b, tag(c), tmp, "xyzw"[c - first]) for c in used))
        if clamp_cols:
            e.open("if (x_edge)")
            for b in range(BY):
                emit_column_fix({c: "rw{}_{}".format(b, tag(c)) for c in cols})
            e.close()
        for c_index, dx in enumerate(dxs):
            for b in range(BY):
                for j in range(4):
                    def tap_text(tap, b=b, j=j, dx=dx):
                        if id(tap) in moving_ids:
                            return "rw{}_{}".format(b, tag(j + dx))
                        tz, ty, tx = _tap3(tap.offset)
                        return name_of(phase, tap.arg, tz, b + ty, j + tx)
                    single = pp.PointProgram(program.dim, program.shape, program.out_ctype, program.args,
                                             [body], program.ghost_depth, program.boundary)
                    current_tap_text[0], current_acc[0] = tap_text, acc[b][j]
                    text = emit_statements(single, tap_text, table_text, acc=acc[b][j], indent="")[0]
                    text = _LOCAL_NAME.sub(lambda m: "{}_{}_{}".format(m.group(0), b, j), text)
                    for slot, param in enumerate(roll['params']):
                        if param is not None:
                            array = "cpar" if param[0] == 'const' else "spar"
                            text = text.replace("@P{}@".format(slot), "{}{}[pk + {}]".format(array, slot, c_index))
                    if guard_halo_body:
                        text = "if (!(halo_mask & {}u)) {{ {} }}".format(1 << (b * 4 + j), text)
                    e(text)
        e.close()

    def emit_release():
        e("__syncwarp();")
        if U == 1:
            e("if (lane == 0 && unit >= {b}) mbar_arrive(&bar_empty[(unit - {b}) % {S}]);".format(b=rel_back, S=S))
        else:
            e.open("if (lane == 0 && unit >= {})".format(rel_back))
            e("const int ur = unit - {};".format(rel_back))
            e("if (ur % {u} == {last} && ur / {u} != unit_hi / {u}) mbar_arrive(&bar_empty[(ur / {u}) % {S}]);".format(
                u=U, last=U - 1, S=S))
            e.close()

    def table_text(ref, index_text):
        constant = _const_index(ref.index)
        if constant is not None:
            return "tc{}_{}".format(ref.arg, constant)
        fast = _fast_floor_index(ref.index)
        if fast is not None and (ref.arg in replicated or ref.arg in smem_tables):
            # the enclosing emitter prints the float operand; see _fast_floor_index
            operand = pp.expr_to_c(fast, current_tap_text[0], table_text, out_text=current_acc[0])
            if ref.arg in replicated:
                # one shift-add forms the shared-memory address: the float's bits are 0x4B000000 + k,
                # entry k of this lane lives 128 k bytes above repk (which has -(0x4B000000 << 7),
                # i.e. 0x80000000 modulo 2^32, and the lane's word folded in)
                return "lds_f32(repk{} + (__float_as_uint(__fadd_rz(fabsf({}), 8388608.0f)) << 7))".format(
                    ref.arg, operand)
            bits = "(__float_as_int(__fadd_rz(fabsf({}), 8388608.0f)) - 0x4B000000)".format(operand)
            return "tab{}[{}]".format(ref.arg, bits)
        if ref.arg in replicated:
            return "rep{}[(({}) << 5) + lane]".format(ref.arg, index_text)
        if ref.arg in smem_tables:
            return "tab{}[{}]".format(ref.arg, index_text)
        return "a{}[{}]".format(ref.arg, index_text)

    if U > 1:
        e("// axis-0 clamping is done here: a unit outside [unit_lo, unit_hi] reads the clamped row")
        e("const int unit_lo = max({} - p_first, 0);".format(c_lo[0]))
        e("const int unit_hi = {} - p_first;".format(c_hi[0]))
    e("int unit = 0;                        // newest streamed unit, relative to p_first")
    e.open("while (true)")
    for phase in range(phases):
        e("// ---- window phase {}".format(phase))
        e("if (unit >= n_units) break;")
        e.open("")
        if U == 1:
            e("mbar_wait(&bar_full[unit % {S}], (unit / {S}) & 1);".format(S=S))
        else:
            e("if (unit % {u} == 0) mbar_wait(&bar_full[(unit / {u}) % {S}], (unit / {us}) & 1);".format(
                u=U, S=S, us=U * S))
        # where each (grid, age) lives in the ring right now
        reads = {}
        for key, ages in fp.ages.items():
            for age in ([max(ages)] if regwin else sorted(ages)):
                reads.setdefault((key[0], age), []).append(key)
        for (g, age) in sorted(reads):
            back = HZ_HI - age
            name = "u{}_{}".format(g, tag(age))
            ring_base = "ring + {}".format(ring_of[g] * S * SF)
            if U == 1:
                e("const float* const {n} = {rb} + ((unit + {k} - {b}) % {S}) * {sf} + sbase;".format(
                    n=name, rb=ring_base, k=S * W, b=back, S=S, sf=SF))
            else:
                e("const int i{n} = min(max(unit - {b}, unit_lo), unit_hi);".format(n=name, b=back))
                e("const float* const {n} = {rb} + ((i{n} / {u}) % {S}) * {sf} + (i{n} % {u}) * {p} + sbase;".format(
                    n=name, rb=ring_base, u=U, S=S, sf=SF, p=PITCH))
        if regwin:
            for (g, age) in sorted(reads):
                emit_loads(phase, g, age, reads[(g, age)])
            emit_release()

        # ---- compute output unit z once the window is full
        e.open("if (unit >= {})".format(W - 1))
        e("const int z = p_first + unit - ({});".format(HZ_HI))
        acc = [["acc{}_{}".format(b, j) for j in range(4)] for b in range(BY)]
        for b in range(BY):
            e("float {};".format(", ".join(n + " = 0.0f" for n in acc[b])))
        loaded = set()
        if not regwin and copy:
            keys = [(0, b, j) for b in range(BY) for j in range(4)]
            emit_loads(phase, 0, 0, keys)
            loaded.update((0, 0, b, j) for b in range(BY) for j in range(4))
        if not clamp:
            ztests = []
            if i_lo[0] > 0:
                ztests.append("z < {}".format(i_lo[0]))
            if i_hi[0] < N0:
                ztests.append("z >= {}".format(i_hi[0]))
            e("const bool z_halo = {};".format(" || ".join(ztests) if ztests else "false"))
            # halo planes/rows are computed like any other and overwritten below: keeping the
            # arithmetic in the same basic block as the window loads lets ptxas overlap them.  Only
            # data-indexed tables (garbage index from zero-filled cells) force a real branch.
            if guard_halo_body:
                e.open("if (!z_halo)")
        for s_index, statement in enumerate(program.statements):
            if roll is not None and roll['start'] <= s_index < roll['stop']:
                if s_index == roll['start']:
                    emit_rolled(phase, loaded, acc)
                continue
            if not regwin:
                wanted = {}
                stack = [statement.expr]
                while stack:
                    node = stack.pop()
                    if isinstance(node, pp.Tap):
                        dz, dy, dx = _tap3(node.offset)
                        rows = {b + dy for b in range(BY)}
                        wanted.setdefault((node.arg, dz), set()).update(rows)
                    stack.extend(node.children())
                for (g, age), rows in sorted(wanted.items()):
                    keys = [k for k, ages in fp.ages.items()
                            if k[0] == g and age in ages and k[1] in rows and (g, age, k[1], k[2]) not in loaded]
                    if keys:
                        emit_loads(phase, g, age, keys)
                        loaded.update((g, age, k[1], k[2]) for k in keys)
            for b in range(BY):
                for j in range(4):
                    def tap_text(tap, b=b, j=j):
                        dz, dy, dx = _tap3(tap.offset)
                        return name_of(phase, tap.arg, dz, b + dy, j + dx)
                    single = pp.PointProgram(program.dim, program.shape, program.out_ctype, program.args,
                                             [statement], program.ghost_depth, program.boundary)
                    current_tap_text[0], current_acc[0] = tap_text, acc[b][j]
                    text = emit_statements(single, tap_text, table_text, acc=acc[b][j], indent="")[0]
                    text = _LOCAL_NAME.sub(lambda m: "{}_{}_{}".format(m.group(0), b, j), text)
                    if guard_halo_body:
                        if isinstance(statement, pp.LocalAssign) and statement.declares:
                            raise StencilException("local temporaries cannot be combined with data-indexed "
                                                   "tables under 'zero'/'copy' boundary handling")
                        text = "if (!(halo_mask & {}u)) {{ {} }}".format(1 << (b * 4 + j), text)
                    e(text)
        if not clamp:
            if guard_halo_body:
                e.close()   # if (!z_halo)
            e.open("if (z_halo || halo_mask)")        # rare: one uniform branch for interior threads
            for b in range(BY):
                for j in range(4):
                    value = name_of(phase, 0, 0, b, j) if copy else "0.0f"
                    e("if (z_halo || (halo_mask & {}u)) {} = {};".format(1 << (b * 4 + j), acc[b][j], value))
            e.close()
        for b in range(BY):
            if rank == 3:
                cond = "col_ok && gy + {} < row_end".format(b)
                address = "out + ((i64)z * {} + (gy + {})) * {} + gx".format(N1, b, domain.row_pitch)
            else:
                cond = "col_ok{}".format(b)
                address = "out + (i64)z * {} + {}".format(domain.row_pitch, col_base(b))
            e.open("if ({})".format(cond))
            e("float* const dst = {};".format(address))
            e("st_global_v4(dst, {});".format(", ".join(acc[b])))
            if domain.is_slab:          # only slabs mirror boundary planes into a neighbour's ghost planes
                e("if (out_peer) st_global_v4(out_peer + (dst - out) + peer_shift, {});".format(", ".join(acc[b])))
            e.close()
        e.close()   # if unit >= W-1
        if not regwin:
            emit_release()
        e("++unit;")
        e.close()
    e.close()   # while
    e.close()   # kernel
    return e.text()


def _fast_floor_index(index):
    """``abs((int)x)`` with x float32 -> x, else None.  For |x| < 2^23 (any valid table index is)
    the value equals floor(|x|), which one round-toward-zero add of 2^23 leaves in the mantissa:
    a full-rate FADD.RZ instead of a quarter-rate F2I plus IABS."""
    if isinstance(index, pp.Call) and index.func == 'abs':
        inner = index.args[0]
        if isinstance(inner, pp.Cast) and inner.ctype == pp.INT and inner.operand.ctype == pp.FLOAT:
            return inner.operand
    return None


def table_plan(program, replicate):
    """(constant-index refs hoisted to registers, tables staged in smem, tables bank-replicated)"""
    data_tables, const_refs = set(), {}
    for node in program.walk():
        if isinstance(node, pp.TableRef):
            constant = _const_index(node.index)
            if constant is None:
                data_tables.add(node.arg)
            else:
                const_refs[(node.arg, constant)] = node.ctype
    candidates = [a for a in program.args if not a.used_as_grid and a.index in data_tables]
    replicated = {a.index: a for a in candidates
                  if replicate and a.ctype == pp.FLOAT and a.size <= 256}
    smem_tables = {a.index: a for a in candidates
                   if a.index not in replicated and a.size * a.dtype.itemsize <= 8192}
    return data_tables, const_refs, smem_tables, replicated


# ---- re-rolling of large neighborhoods -------------------------------------------------------------
class _Pure(object):
    """placeholder for a tap-dependent but data-independent sub-expression (a folded distance(),
    a table entry at a folded index): becomes one entry per tap of a parameter table"""
    def __init__(self, slot, ctype):
        self.slot, self.ctype = slot, ctype


def _is_pure(expr):
    if isinstance(expr, (pp.Tap, pp.OutRef, pp.LocalRef)):
        return False
    if isinstance(expr, pp.TableRef) and _const_index(expr.index) is None:
        return False
    return all(_is_pure(c) for c in expr.children())


def _shape_of(expr, taps, pures):
    """structural key of an expression with taps and maximal pure sub-expressions abstracted"""
    if isinstance(expr, pp.Tap):
        taps.append(expr)
        return ('tap', expr.arg, expr.ctype)
    if _is_pure(expr):
        pures.append(expr)
        return ('pure', expr.ctype)
    if isinstance(expr, pp.OutRef):
        return ('out',)
    if isinstance(expr, pp.LocalRef):
        return ('local', expr.name)
    head = (type(expr).__name__, getattr(expr, 'op', None), getattr(expr, 'func', None), expr.ctype,
            getattr(expr, 'arg', None))
    return head + tuple(_shape_of(c, taps, pures) for c in expr.children())


def _pure_value(expr):
    """('const', value) or ('table', arg, index) for the supported pure forms, else None"""
    if isinstance(expr, pp.Const):
        return ('const', expr.value)
    if isinstance(expr, pp.TableRef):
        return ('table', expr.arg, _const_index(expr.index))
    return None


def find_rolled_run(program, min_taps=48):
    """Longest run of consecutive stores that are one statement template applied over a full
    (dy, dx) rectangle of tap offsets in row-major order.  Returns None if there is none worth
    rolling.  Rolling keeps the statement order, hence the bits of the result."""
    statements = program.statements
    infos = []
    for statement in statements:
        if not isinstance(statement, pp.Store):
            infos.append(None)
            continue
        taps, pures = [], []
        key = (statement.op, _shape_of(statement.expr, taps, pures))
        values = [_pure_value(q) for q in pures]
        infos.append(None if any(v is None for v in values) else (key, taps, pures, values))
    best = None
    start = 0
    while start < len(statements):
        if infos[start] is None:
            start += 1
            continue
        stop = start
        while stop < len(statements) and infos[stop] is not None and infos[stop][0] == infos[start][0]:
            stop += 1
        if stop - start >= min_taps and (best is None or stop - start > best[1] - best[0]):
            best = (start, stop)
        start = stop
    if best is None:
        return None
    start, stop = best
    run = infos[start:stop]
    n_taps = len(run[0][1])
    moving = [k for k in range(n_taps) if len({info[1][k].offset for info in run}) > 1]
    if not moving:
        return None
    offsets = []
    for info in run:
        these = {info[1][k].offset for k in moving}
        if len(these) != 1:
            return None                                   # several different moving offsets in one statement
        offsets.append(_tap3(next(iter(these))))
    if len({o[0] for o in offsets}) != 1 or len({info[1][moving[0]].arg for info in run}) != 1:
        return None
    dys = sorted({o[1] for o in offsets})
    dxs = sorted({o[2] for o in offsets})
    expected = [(offsets[0][0], dy, dx) for dy in dys for dx in dxs]
    if offsets != expected or dys != list(range(dys[0], dys[-1] + 1)) or dxs != list(range(dxs[0], dxs[-1] + 1)):
        return None
    if len(dys) < 3:
        return None
    params = []
    for slot in range(len(run[0][2])):
        values = [info[3][slot] for info in run]
        kinds = {v[0] for v in values}
        if len(kinds) != 1:
            return None
        if len(set(values)) == 1:
            params.append(None)                           # same for every tap: stays inline
        elif kinds == {'const'}:
            params.append(('const', run[0][2][slot].ctype, [v[1] for v in values]))
        else:
            if len({v[1] for v in values}) != 1:
                return None
            params.append(('table', run[0][2][slot].ctype, values[0][1], [v[2] for v in values]))
    return dict(start=start, stop=stop, dz=offsets[0][0], dys=dys, dxs=dxs, moving=set(moving),
                params=params, template=statements[start], grid=run[0][1][moving[0]].arg)


def _const_index(index):
    if isinstance(index, pp.Const):
        return int(index.value)
    if isinstance(index, pp.Cast) and isinstance(index.operand, pp.Const):
        return int(index.operand.value)
    return None


def make_plan(program, domain, device_props, options, override=None):
    from .transformer import CudaPlan
    from .planner import Launch
    from .domain import Domain
    original = program
    embedded = wants_plane_embedding(program, domain)
    if embedded:
        program = embed_as_plane(program)
        domain3 = Domain(program.shape, program.ghost_depth, row_pitch=domain.row_pitch)
        cfg = choose_config(program, domain3, device_props, dict(override or {}, embedded=True))
        source = generate(program, domain3, cfg)
    else:
        cfg = choose_config(program, domain, device_props, override)
        source = generate(program, domain, cfg)
    rank = cfg.rank
    shape = program.shape
    row_pitch = domain.row_pitch             # elements between row starts (== shape[-1] unless padded)
    specs = []
    for arg in _grid_args(program):
        if rank == 3:
            n0, n1, n2 = shape
            specs.append(TensorMapSpec(arg.index, (n2, n1, n0), (row_pitch * 4, n1 * row_pitch * 4),
                                       (cfg.PITCH, cfg.ROWS, 1), tag="tile"))
        else:
            n0, n1 = shape
            specs.append(TensorMapSpec(arg.index, (n1, n0), (row_pitch * 4,), (cfg.PITCH, cfg.U), tag="block"))
    tiles_x = -(-shape[-1] // cfg.TX)
    tiles_y = -(-shape[1] // cfg.TY) if rank == 3 else 1
    slots = device_props.sm_count * cfg.blocks_per_sm
    min_chunk = cfg.min_chunk
    static_smem = sum(a.size * 128 for a in table_plan(program, embedded)[3].values())

    def geometry(a0_begin, a0_end):
        if embedded:
            rows = a0_end - a0_begin
            launch = Launch((tiles_x, -(-rows // cfg.TY), 1), (cfg.threads, 1, 1), cfg.smem_bytes)
            launch.extra = (1,)
            return launch
        n_units = a0_end - a0_begin
        chunks = plan_chunks_adaptive(n_units, tiles_x * tiles_y, slots, cfg.W, min_chunk, device_props.sm_count)
        chunk_len = -(-n_units // chunks)
        if rank == 2:
            chunk_len = -(-chunk_len // cfg.U) * cfg.U if chunk_len > cfg.U else chunk_len
        chunks = -(-n_units // chunk_len)
        launch = Launch((tiles_x, tiles_y, chunks), (cfg.threads, 1, 1), cfg.smem_bytes)
        launch.extra = (chunk_len,)
        return launch

    notes = dict(cfg.describe(), embedded_plane=embedded, static_smem=static_smem)
    return CudaPlan(original, domain, 'stream', source, FUNCTION, options, geometry,
                    tensor_maps=specs, notes={'config': notes})
