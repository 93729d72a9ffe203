"""Iterated sweeps on small grids: per-sweep time with the native ping-pong launch loop."""
import time
import numpy, torch
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.jacobi_stencil import Jacobi

for name, stencil, shape in (("jacobi3d", Jacobi3D(1.0, 0.25, backend='cuda'), (64, 64, 64)),
                             ("jacobi3d", Jacobi3D(1.0, 0.25, backend='cuda'), (128, 128, 128)),
                             ("jacobi3d", Jacobi3D(1.0, 0.25, backend='cuda'), (256, 256, 256)),
                             ("jacobi2d", Jacobi(backend='cuda'), (1024, 1024))):
    grid = torch.rand(shape, device='cuda')
    for n in (10, 1000):
        stencil.iterate(grid, n); torch.cuda.synchronize()
        t = time.perf_counter(); out = stencil.iterate(grid, n); torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print("{} {} n={} : {:.2f} us/sweep  {:.1f} GStencil/s".format(
            name, shape, n, dt / n * 1e6, numpy.prod(shape) * n / dt / 1e9), flush=True)
