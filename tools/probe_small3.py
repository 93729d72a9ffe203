"""(1) is the iterated loop host-bound?  (2) rank-2 small grids: stream tile / min_chunk sweep."""
import os, sys, time, subprocess

CHILD = r'''
import os, sys, time, numpy, torch
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.jacobi_stencil import Jacobi
shape = tuple(int(v) for v in sys.argv[1].split("x"))
stencil = Jacobi3D(1.0, 0.25, backend='cuda') if len(shape) == 3 else Jacobi(backend='cuda')
grid = torch.rand(shape, device='cuda')
n = 1000
stencil.iterate(grid, n); torch.cuda.synchronize()
best = 1e9; best_host = 1e9
for _ in range(3):
    t = time.perf_counter(); stencil.iterate(grid, n); h = time.perf_counter() - t; torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t); best_host = min(best_host, h)
concrete = stencil.specializer.specialize((grid,))
geo = concrete.plan.geometry(*concrete.plan.domain.sweep_range)
print("{:.2f} us/sweep (host enqueue {:.2f}) {:.0f} GSt/s plan={} grid={}".format(best / n * 1e6, best_host / n * 1e6,
      numpy.prod(shape) * n / best / 1e9, concrete.plan.strategy, geo.grid))
'''

def run(shape, **env):
    e = dict(os.environ, PYTHONPATH=".", **env)
    out = subprocess.run([sys.executable, "-c", CHILD, shape], env=e, capture_output=True, text=True)
    text = (out.stdout.strip().splitlines() or [out.stderr.strip().splitlines()[-1] if out.stderr.strip() else "?"])[-1]
    print(shape, env.get("STENCIL_CUDA_STREAM", env.get("STENCIL_CUDA_STRATEGY")), "->", text, flush=True)

run("32x32x32", STENCIL_CUDA_STRATEGY="direct")
run("64x64x64", STENCIL_CUDA_STRATEGY="direct")
run("64x64x64", STENCIL_CUDA_STRATEGY="stream", STENCIL_CUDA_TEMPORAL="0", STENCIL_CUDA_STREAM="TY=16,BY=4,NCW=4,prefetch=3,min_chunk=4")
run("128x128x128", STENCIL_CUDA_STRATEGY="stream", STENCIL_CUDA_TEMPORAL="0", STENCIL_CUDA_STREAM="TY=16,BY=4,NCW=4,prefetch=3,min_chunk=4")
run("128x128x128", STENCIL_CUDA_STRATEGY="stream", STENCIL_CUDA_TEMPORAL="0", STENCIL_CUDA_STREAM="TY=16,BY=4,NCW=4,prefetch=3,min_chunk=2")
run("128x128x128", STENCIL_CUDA_STRATEGY="stream", STENCIL_CUDA_TEMPORAL="0", STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=3,min_chunk=2")
for shape in ("512x512", "1024x1024", "2048x2048", "4096x4096"):
    run(shape, STENCIL_CUDA_STRATEGY="direct")
    for ncw, u, by in ((4, 4, 2), (4, 4, 1), (2, 4, 1), (1, 4, 1), (2, 8, 1)):
        for mc in (8, 16, 64):
            run(shape, STENCIL_CUDA_STRATEGY="stream",
                STENCIL_CUDA_STREAM="NCW={},U={},BY={},min_chunk={}".format(ncw, u, by, mc))
