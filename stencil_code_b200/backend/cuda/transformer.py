"""StencilCudaTransformer -- the ``backend_dict['cuda']`` entry.

Takes the place of reference ``backend/ocl.py:55-618`` (StencilOclTransformer) in the plug-in
slot described in SURVEY.md section 8b: same constructor signature as every reference backend
(``backend/stencil_backend.py:12-13``), same kernel attributes consumed, mandatory
``visit_InteriorPointsLoop``.  Instead of an OpenCL kernel plus a C host function it produces a
``CudaPlan``: the typed per-point program, the iteration domain, the strategy the planner chose,
generated sm_100a CUDA source and launch geometry.  Nothing here touches a device.
"""
import os

from ...stencil_exception import StencilException
from .. import point_program as pp
from ..stencil_backend import StencilBackend
from .domain import Domain
from . import planner


class CudaPlan(object):
    """Everything needed to compile and launch one specialised stencil."""

    def __init__(self, program, domain, strategy, source, function, options, geometry,
                 tensor_maps=(), notes=None):
        self.program = program
        self.domain = domain
        self.strategy = strategy              # 'direct' | 'stream'
        self.source = source
        self.function = function
        self.options = tuple(options)
        self.geometry = geometry              # callable (a0_begin, a0_end) -> Launch
        self.tensor_maps = tuple(tensor_maps)  # TensorMapSpec per TMA-staged grid argument
        self.notes = notes or {}

    def describe(self):
        return "{} strategy, {} grid {}, boundary {}".format(
            self.strategy, self.program.out_ctype, self.program.shape, self.program.boundary)


def strict_double_requested():
    """STENCIL_CUDA_STRICT=1 keeps every double sub-expression in double (no exact demotion)."""
    return os.environ.get("STENCIL_CUDA_STRICT", "0") not in ("", "0")


NVRTC_OPTIONS = (
    "--std=c++17",
    "--fmad=false",        # the reference's C is compiled without FMA contraction (-std=c99, x86-64)
    "--prec-div=true",
    "--prec-sqrt=true",
    "--ftz=false",
    "-lineinfo",
)


class StencilCudaTransformer(StencilBackend):
    def __init__(self, input_grids=None, output_grid=None, kernel=None, arg_cfg=None,
                 fusable_nodes=None, testing=False, open_faces=(False, False), device_props=None,
                 strategy=None, slab_ghost=None, row_pitch=None):
        super(StencilCudaTransformer, self).__init__(
            input_grids, output_grid, kernel, arg_cfg=arg_cfg, fusable_nodes=fusable_nodes,
            testing=testing)
        self.open_faces = open_faces
        self.slab_ghost = slab_ghost
        self.row_pitch = row_pitch
        self.device_props = device_props
        self.forced_strategy = strategy or os.environ.get("STENCIL_CUDA_STRATEGY") or None

    def build_program(self, node):
        statements = self.lower_interior_body(node)
        if not any(isinstance(s, pp.Store) for s in statements):
            raise StencilException("the kernel never assigns the output grid")
        shape = tuple(self.arg_cfg[0].shape)
        if len(shape) != self.kernel.dim:
            raise StencilException(
                "grid has {} dimensions but the stencil's neighborhoods have {}".format(
                    len(shape), self.kernel.dim))
        if self.out_ctype not in (pp.FLOAT, pp.DOUBLE):
            raise StencilException("the first argument (the grid) must be float32 or float64")
        # 'wrap' is accepted and, exactly as in the reference, behaves like 'zero'
        boundary = {'clamp': 'clamp', 'copy': 'copy', 'periodic': 'periodic'}.get(
            self.kernel.boundary_handling, 'zero')
        if boundary == 'periodic':
            if any(self.open_faces):
                raise StencilException("periodic boundaries are not available on slab decompositions")
            if any(g > n for g, n in zip(self.kernel.ghost_depth, shape)):
                raise StencilException("periodic boundaries need a grid at least as wide as the stencil's reach")
        args = [self.input_dict[name] for name in self.input_names]
        args[0].used_as_grid = True
        program = pp.PointProgram(self.kernel.dim, shape, self.out_ctype, args, statements,
                                  self.kernel.ghost_depth, boundary)
        lo, hi = program.tap_extent()
        for d in range(program.dim):
            if -lo[d] > program.ghost_depth[d] or hi[d] > program.ghost_depth[d]:
                raise StencilException("tap offsets exceed the ghost depth")   # cannot happen
        if not strict_double_requested():
            program.demoted = pp.demote_exact(program)
        else:
            program.demoted = 0
        return program

    def visit_InteriorPointsLoop(self, node):
        program = self.build_program(node)
        domain = Domain(program.shape, program.ghost_depth, *self.open_faces, slab_ghost=self.slab_ghost,
                        row_pitch=self.row_pitch)
        for d in range(program.dim):
            if domain.interior_hi[d] <= domain.interior_lo[d] and program.boundary != 'clamp' \
                    and program.shape[d] < 2 * program.ghost_depth[d]:
                # reference loops would simply not execute; nothing special to do
                pass
        return planner.plan(program, domain, device_props=self.device_props,
                            forced_strategy=self.forced_strategy, options=NVRTC_OPTIONS)
