#!/usr/bin/env python
"""Iterated-sweep timing: python tools/probe_iter.py SHAPE SWEEPS [BOUNDARY]  (Jacobi3D)"""
import ctypes, os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stencil_code_b200.backend.cuda import runtime
from stencil_code_b200.library.jacobi_3d import Jacobi3D

shape = tuple(int(v) for v in sys.argv[1].split("x"))
sweeps = int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "clamp"
stencil = Jacobi3D(1.0, 0.25, boundary_handling=mode)
device = runtime.Device.get(0)
library = runtime.load_library()
points = float(numpy.prod(shape))
a, b = device.alloc(int(points) * 4), device.alloc(int(points) * 4)
geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
runtime.launch(device.utils(), "sc_fill_uniform", geometry,
               [ctypes.c_void_p(a.ptr), ctypes.c_uint64(int(points)), ctypes.c_uint64(1), ctypes.c_uint64(0),
                ctypes.c_float(1.0)], device.stream)


class Fake(object):
    def __init__(self, shape):
        self.shape, self.dtype = shape, numpy.dtype('float32')
concrete = stencil.specializer.specialize((Fake(shape),))
concrete.sweeps(a.ptr, b.ptr, [], 4, device.stream)
device.synchronize()
start, stop = device.new_event(True), device.new_event(True)
library.sc_event_record(start, device.stream)
concrete.sweeps(a.ptr, b.ptr, [], sweeps, device.stream)
library.sc_event_record(stop, device.stream)
device.synchronize()
elapsed = ctypes.c_float()
library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed))
fused = concrete.temporal()
print("jacobi3d {} {} x{} sweeps: {:.3f} ms/sweep  {:.1f} GStencil/s  ({}) {}".format(
    "x".join(map(str, shape)), mode, sweeps, elapsed.value / sweeps, points * sweeps / elapsed.value / 1e6,
    "temporal " + str(fused[0].notes['config']) if fused else "plain sweeps",
    {k: v for k, v in os.environ.items() if k.startswith("STENCIL_CUDA")}))
