"""End-to-end (host arrays in/out) C2 call: pipeline chunk size sweep, and the raw PCIe copy rates."""
import time, ctypes
import numpy, torch
from stencil_code_b200 import pinned_empty
from stencil_code_b200.backend.cuda import runtime
from stencil_code_b200.library.laplacian import LaplacianKernel

shape = (512, 512, 512)
grid = pinned_empty(shape); grid[...] = numpy.random.RandomState(1).random_sample(shape).astype(numpy.float32)
out = pinned_empty(shape)
stencil = LaplacianKernel(backend='cuda')
stencil(grid, out=out)
concrete = stencil.specializer.specialize((grid,))
device = concrete.device
lib = runtime.load_library()

# raw copies
buf = device.alloc(grid.nbytes); buf2 = device.alloc(grid.nbytes)
def timed(fn, n=5):
    fn(); device.synchronize(device.stream); device.synchronize(device.h2d_stream); device.synchronize(device.d2h_stream)
    best = 1e9
    for _ in range(n):
        t = time.perf_counter(); fn()
        device.synchronize(device.stream); device.synchronize(device.h2d_stream); device.synchronize(device.d2h_stream)
        best = min(best, time.perf_counter() - t)
    return best
h2d = lambda: lib.sc_memcpy_h2d_async(ctypes.c_void_p(buf.ptr), ctypes.c_void_p(grid.ctypes.data), grid.nbytes, device.h2d_stream)
d2h = lambda: lib.sc_memcpy_d2h_async(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(buf2.ptr), grid.nbytes, device.d2h_stream)
both = lambda: (h2d(), d2h())
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both directions at once", both)):
    t = timed(fn)
    print("{}: {:.2f} ms  {:.1f} GB/s per direction".format(name, t * 1e3, grid.nbytes / t / 1e9), flush=True)
buf.release(); buf2.release()

for mib in (4, 8, 16, 32, 64):
    runtime.ConcreteCudaStencil.PIPELINE_CHUNK_BYTES = mib << 20
    t = timed(lambda: stencil(grid, out=out), n=7)
    print("chunk {} MiB: {:.2f} ms  {:.2f} GStencil/s".format(mib, t * 1e3, grid.size / t / 1e9), flush=True)
