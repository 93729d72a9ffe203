#!/usr/bin/env python
"""Iterated-sweep timing on rank-2 grids: python tools/probe_iter2d.py KERNEL SHAPE SWEEPS [BOUNDARY]
KERNEL: jacobi2d | simple | conv5"""
import ctypes, os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stencil_code_b200.backend.cuda import runtime
from stencil_code_b200.library.jacobi_stencil import Jacobi
from stencil_code_b200.library.simple import SimpleKernel
from stencil_code_b200.library.convolution import ConvolutionFilter

name = sys.argv[1]
shape = tuple(int(v) for v in sys.argv[2].split("x"))
sweeps = int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "clamp"
kernel = numpy.ones((5, 5)) / 32.0
stencil = {"jacobi2d": lambda: Jacobi(boundary_handling=mode), "simple": lambda: SimpleKernel(boundary_handling=mode),
           "conv5": lambda: ConvolutionFilter(convolution_array=kernel, boundary_handling=mode)}[name]()
device = runtime.Device.get(0)
library = runtime.load_library()
points = float(numpy.prod(shape))
a, b = device.alloc(int(points) * 4), device.alloc(int(points) * 4)
geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
runtime.launch(device.utils(), "sc_fill_uniform", geometry,
               [ctypes.c_void_p(a.ptr), ctypes.c_uint64(int(points)), ctypes.c_uint64(1), ctypes.c_uint64(0),
                ctypes.c_float(1e-3)], device.stream)


class Fake(object):
    def __init__(self, shape, dtype='float32'):
        self.shape, self.dtype = shape, numpy.dtype(dtype)
tables, table_pointers = (), []
if name == "conv5":
    tables = (Fake(stencil.coefficients.shape, stencil.coefficients.dtype),)
    t = device.alloc(stencil.coefficients.nbytes)
    table_pointers = [t.ptr]
concrete = stencil.specializer.specialize((Fake(shape),) + tables)
concrete.sweeps(a.ptr, b.ptr, table_pointers, 4, device.stream)
device.synchronize()
start, stop = device.new_event(True), device.new_event(True)
library.sc_event_record(start, device.stream)
concrete.sweeps(a.ptr, b.ptr, table_pointers, sweeps, device.stream)
library.sc_event_record(stop, device.stream)
device.synchronize()
elapsed = ctypes.c_float()
library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed))
fused = concrete.temporal()
print("{} {} {} x{} sweeps: {:.3f} ms/sweep  {:.1f} GStencil/s  ({}) {}".format(
    name, "x".join(map(str, shape)), mode, sweeps, elapsed.value / sweeps, points * sweeps / elapsed.value / 1e6,
    "temporal " + str(fused[0].notes['config']) if fused else "plain sweeps " + str(concrete.plan.notes.get('config')),
    {k: v for k, v in os.environ.items() if k.startswith("STENCIL_CUDA")}))
