"""Small/medium grids, iterated: us/sweep for every strategy / tile choice."""
import os, sys, time, subprocess
import numpy

CHILD = r'''
import os, sys, time, numpy, torch
from stencil_code_b200.library.jacobi_3d import Jacobi3D
shape = tuple(int(v) for v in sys.argv[1].split("x"))
stencil = Jacobi3D(1.0, 0.25, backend='cuda')
grid = torch.rand(shape, device='cuda')
n = 400
stencil.iterate(grid, n); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    t = time.perf_counter(); stencil.iterate(grid, n); torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t)
concrete = stencil.specializer.specialize((grid,))
fused = concrete.temporal()
geo = concrete.plan.geometry(*concrete.plan.domain.sweep_range)
print("{:.2f} us/sweep {:.0f} GSt/s plan={} grid={} fused={}".format(best / n * 1e6, numpy.prod(shape) * n / best / 1e9,
      concrete.plan.strategy, geo.grid, None if fused is None else getattr(fused[0], "description", "yes")))
'''

def run(shape, **env):
    e = dict(os.environ, PYTHONPATH=".", **env)
    out = subprocess.run([sys.executable, "-c", CHILD, shape], env=e, capture_output=True, text=True)
    text = (out.stdout.strip().splitlines() or [out.stderr.strip().splitlines()[-1] if out.stderr.strip() else "?"])[-1]
    print(shape, env.get("STENCIL_CUDA_STREAM", env.get("STENCIL_CUDA_STRATEGY")), "->", text, flush=True)

for shape in ("64x64x64", "128x128x128", "192x192x192", "256x256x256"):
    run(shape, STENCIL_CUDA_STRATEGY="direct")
    for cfg in ("32,4,8,3", "16,4,4,3", "8,2,4,3", "8,1,8,3"):
        for mc in (4, 8, 16):
            ty, by, ncw, pf = cfg.split(",")
            run(shape, STENCIL_CUDA_STRATEGY="stream", STENCIL_CUDA_TEMPORAL="0",
                STENCIL_CUDA_STREAM="TY={},BY={},NCW={},prefetch={},min_chunk={}".format(ty, by, ncw, pf, mc))
