"""Default plans, iterated: us/sweep with and without programmatic dependent launch."""
import os, sys, subprocess

CHILD = r'''
import os, sys, time, numpy, torch
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.jacobi_stencil import Jacobi
shape = tuple(int(v) for v in sys.argv[1].split("x"))
stencil = Jacobi3D(1.0, 0.25, backend='cuda') if len(shape) == 3 else Jacobi(backend='cuda')
grid = torch.rand(shape, device='cuda')
n = int(sys.argv[2])
stencil.iterate(grid, n); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    t = time.perf_counter(); stencil.iterate(grid, n); torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t)
concrete = stencil.specializer.specialize((grid,))
geo = concrete.plan.geometry(*concrete.plan.domain.sweep_range)
print("{:.2f} us/sweep {:.0f} GSt/s plan={} grid={} fused={}".format(best / n * 1e6,
      numpy.prod(shape) * n / best / 1e9, concrete.plan.strategy, geo.grid, concrete.temporal() is not None))
'''

def run(shape, n, **env):
    e = dict(os.environ, PYTHONPATH=".", **env)
    out = subprocess.run([sys.executable, "-c", CHILD, shape, str(n)], env=e, capture_output=True, text=True)
    text = (out.stdout.strip().splitlines() or [out.stderr.strip().splitlines()[-1] if out.stderr.strip() else "?"])[-1]
    print(shape, env, "->", text, flush=True)

import sys as _sys
if len(_sys.argv) > 1 and _sys.argv[1] == "odd":
    for shape, n in (("513x513x513", 40), ("257x1025x1025", 40), ("8193x8193", 40), ("1025x1025", 400)):
        run(shape, n, STENCIL_CUDA_PITCHED="0")
        run(shape, n, STENCIL_CUDA_PITCHED="1")
    _sys.exit(0)
for shape, n in (("32x32x32", 1000), ("64x64x64", 1000), ("128x128x128", 1000), ("192x192x192", 400), ("256x256x256", 400),
                 ("512x512x512", 100), ("512x1024x1024", 40), ("512x512", 1000), ("1024x1024", 1000),
                 ("2048x2048", 1000), ("4096x4096", 400), ("16384x16384", 20)):
    run(shape, n, STENCIL_CUDA_PDL="0")
    run(shape, n, STENCIL_CUDA_PDL="1")
