#!/usr/bin/env python
"""Launch-gap probe: per-launch time of the C2 kernel for 1, 2, 5, 20 back-to-back launches."""
import ctypes, os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from stencil_code_b200.backend.cuda import runtime

workload = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "laplacian512", None, "cuda")
shape = workload["shape"]
device = runtime.Device.get(0)
library = runtime.load_library()
points = int(numpy.prod(shape))
a, b = device.alloc(points * 4), device.alloc(points * 4)
geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
runtime.launch(device.utils(), "sc_fill_uniform", geometry,
               [ctypes.c_void_p(a.ptr), ctypes.c_uint64(points), ctypes.c_uint64(1), ctypes.c_uint64(0), ctypes.c_float(1.0)],
               device.stream)
concrete = workload["stencil"].specializer.specialize((bench._Shape(shape),))
for _ in range(5):
    concrete.launch([a.ptr], b.ptr)
device.synchronize()
for count in (1, 1, 2, 5, 20, 100):
    start, stop = device.new_event(True), device.new_event(True)
    library.sc_event_record(start, device.stream)
    for _ in range(count):
        concrete.launch([a.ptr], b.ptr)
    library.sc_event_record(stop, device.stream)
    device.synchronize()
    elapsed = ctypes.c_float()
    library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed))
    print("{:4d} launches: {:.2f} us per launch".format(count, elapsed.value * 1e3 / count))
