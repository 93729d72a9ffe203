"""Print the plan the transformer picks for a list of shapes (CPU only)."""
import sys
import numpy
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.jacobi_stencil import Jacobi
from stencil_code_b200.library.laplacian import LaplacianKernel
from stencil_code_b200.library.convolution import ConvolutionFilter

class FakeArray(object):
    def __init__(self, shape):
        self.shape, self.dtype, self.ndim = shape, numpy.dtype('float32'), len(shape)

for shape in [(64,)*3, (96,)*3, (128,)*3, (192,)*3, (256,)*3, (512,)*3, (2048, 1024, 1024), (256, 1024, 1024),
              (512, 512), (1024, 1024), (2048, 2048), (4096, 4096), (16384, 16384)]:
    stencil = Jacobi3D(1.0, 0.25, backend='cuda') if len(shape) == 3 else Jacobi(backend='cuda')
    spec = stencil.specializer
    plan = spec.transform(spec.args_to_subconfig((FakeArray(shape),)), (False, False), None)
    geo = plan.geometry(*plan.domain.sweep_range)
    cfg = (plan.notes or {}).get('config', {})
    print(shape, plan.strategy, {k: cfg.get(k) for k in ('TY', 'BY', 'NCW', 'U', 'S', 'blocks_per_sm')}, geo.grid, getattr(geo, 'extra', None))
