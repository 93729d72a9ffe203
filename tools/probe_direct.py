"""The point kernel (strategy 'direct'), flat vs tiled: GStencil/s on grids the TMA path does not take."""
import os, sys, subprocess

CHILD = r'''
import sys, time, numpy, torch
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.convolution import ConvolutionFilter
from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
kind, mode, dtype = sys.argv[1:4]
dtype = getattr(torch, dtype)
if kind == "jacobi3d":
    stencil, shape, tables = Jacobi3D(1.0, 0.25, backend='cuda', boundary_handling=mode), (513, 512, 513), ()
elif kind == "lap27":
    stencil, shape, tables = SpecializedLaplacian27(backend='cuda', boundary_handling=mode), (257, 512, 513), (torch.tensor([1.0, 0.5, 0.25, 0.125], device='cuda'),)
else:
    kernel = numpy.ones((5, 5), numpy.int64); kernel[2, 2] = -4
    stencil, shape, tables = ConvolutionFilter(kernel, backend='cuda', boundary_handling=mode), (8192, 8193), ()
grid = torch.rand(shape, device='cuda', dtype=dtype)
stencil(grid, *tables); torch.cuda.synchronize()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for _ in range(5):
    stencil(grid, *tables)
stop.record(); torch.cuda.synchronize()
ms = start.elapsed_time(stop) / 5
concrete = stencil.specializer.specialize((grid,) + tables + ((stencil.coefficients,) if kind == "conv5" else ()))
print("{:.3f} ms {:.0f} GStencil/s {} {}".format(ms, numpy.prod(shape) / ms / 1e6, concrete.plan.strategy, concrete.plan.notes.get('config')))
'''
if len(sys.argv) > 1 and sys.argv[1] == "tiles":
    for kind, mode, dtype, tiles in (("jacobi3d", "clamp", "float32", ("4,2", "8,2", "4,4", "8,1", "2,2", "2,4")),
                                     ("jacobi3d", "clamp", "float64", ("4,1", "4,2", "8,1", "2,2")),
                                     ("conv5", "clamp", "float32", ("4,1", "8,1", "2,1")),
                                     ("lap27", "clamp", "float32", ("4,1", "4,2", "2,2", "8,1"))):
        for tile in tiles:
            env = dict(os.environ, PYTHONPATH=".", STENCIL_CUDA_DIRECT_TILE=tile)
            out = subprocess.run([sys.executable, "-c", CHILD, kind, mode, dtype], env=env, capture_output=True, text=True)
            text = (out.stdout.strip().splitlines() or [(out.stderr.strip().splitlines() or ["?"])[-1]])[-1]
            print(kind, mode, dtype, tile, "->", text, flush=True)
    sys.exit(0)
for kind, mode, dtype in (("jacobi3d", "clamp", "float32"), ("jacobi3d", "zero", "float32"), ("jacobi3d", "periodic", "float32"),
                          ("jacobi3d", "clamp", "float64"), ("lap27", "clamp", "float32"), ("conv5", "clamp", "float32"),
                          ("conv5", "periodic", "float32")):
    for variant in ("flat", "tiled"):
        env = dict(os.environ, PYTHONPATH=".", STENCIL_CUDA_DIRECT=variant)
        out = subprocess.run([sys.executable, "-c", CHILD, kind, mode, dtype], env=env, capture_output=True, text=True)
        text = (out.stdout.strip().splitlines() or [(out.stderr.strip().splitlines() or ["?"])[-1]])[-1]
        print(kind, mode, dtype, variant, "->", text, flush=True)
