"""Tiling / launch-geometry planner of the CUDA backend.

Replaces reference ``backend/local_size_computer.py:17-253`` (the OpenCL work-group shape
chooser); the per-dimension halo kernels of ``backend/ocl_boundary_copier.py:69-84`` have no
counterpart: every generated kernel handles its own boundary points.  Like the
reference's planner it is pure host arithmetic over a device-properties record, so it can be
unit-tested with a fake device (the reference does that with ``MockDevice``,
``test/fixtures.py:3-11``); unlike it, it chooses between code-generation strategies:

  direct  one thread per point, global loads (``codegen_direct``) -- any shape / dtype / rank
  stream  TMA-staged shared-memory tiles streamed along axis 0 with a register window
          (``codegen_stream``) -- float32 grids of rank 2 or 3 whose rows are 16-byte multiples
"""
from ...stencil_exception import StencilException
from . import codegen_direct


class DeviceProps(object):
    """The few device facts the planner uses; defaults describe a B200 (sm_100a)."""
    def __init__(self, sm_count=148, smem_per_sm=233472, smem_per_block_optin=232448,
                 regs_per_sm=65536, max_threads_per_sm=2048, l2_bytes=126 * 1024 * 1024,
                 name="NVIDIA B200"):
        self.sm_count = sm_count
        self.smem_per_sm = smem_per_sm
        self.smem_per_block_optin = smem_per_block_optin
        self.regs_per_sm = regs_per_sm
        self.max_threads_per_sm = max_threads_per_sm
        self.l2_bytes = l2_bytes
        self.name = name


B200 = DeviceProps()


class Launch(object):
    def __init__(self, grid, block, smem=0, cluster=(1, 1, 1)):
        self.grid, self.block, self.smem, self.cluster = tuple(grid), tuple(block), int(smem), tuple(cluster)


def plan(program, domain, device_props=None, forced_strategy=None, options=()):
    from .transformer import CudaPlan
    device_props = device_props or B200
    strategy = forced_strategy
    reason = "forced" if forced_strategy else None
    if domain.is_pitched:
        strategy = 'stream'              # padded rows exist only for the TMA kernels (iterate())
    from . import codegen_stream
    rows_kernel = False
    if program.dim == 2 and not codegen_stream.wants_plane_embedding(program, domain):
        from . import codegen_temporal2d
        rows_kernel = codegen_temporal2d.single_applicable(program, domain)
    if strategy is None:
        ok, reason = codegen_stream.applicable(program, domain, device_props)
        if not ok and rows_kernel and program.npoints >= (1 << 20):
            ok, reason = True, "rank-2 {} grid on the row pipeline".format(program.out_ctype)   # float64
        strategy = 'stream' if ok else 'direct'
    if strategy == 'stream':
        if rows_kernel:
            # rank-2 single sweeps: the row pipeline with compile-time ring indices (float32 / float64)
            return codegen_temporal2d.make_single_plan(program, domain, device_props, options)
        ok, why = codegen_stream.applicable(program, domain, device_props, forced=True)
        if not ok:
            raise StencilException("strategy 'stream' cannot handle this stencil: " + why)
        return codegen_stream.make_plan(program, domain, device_props, options)
    if strategy != 'direct':
        raise StencilException("unknown strategy '{}'".format(strategy))
    config = codegen_direct.direct_config(program, domain)
    source = codegen_direct.generate(program, domain, config)

    def geometry(a0_begin, a0_end):
        grid, block = codegen_direct.launch_geometry(program, a0_begin, a0_end, config)
        return Launch(grid, block)

    return CudaPlan(program, domain, 'direct', source, 'stencil_direct', options, geometry,
                    notes={'why': reason, 'config': config})
