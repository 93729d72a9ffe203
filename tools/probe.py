#!/usr/bin/env python
"""Kernel-only timing probe for tuning: python tools/probe.py WORKLOAD [BOUNDARY] [STEPS]
Environment knobs: STENCIL_CUDA_STREAM, STENCIL_CUDA_L2_PROMOTION, STENCIL_CUDA_STRATEGY."""
import ctypes
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from stencil_code_b200.backend.cuda import runtime  # noqa: E402


def main():
    name = sys.argv[1]
    boundary = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    shape = None
    if len(sys.argv) > 4:
        shape = tuple(int(v) for v in sys.argv[4].split("x"))
    workload = bench.make_workload(name, boundary, "cuda")
    shape = shape or workload["shape"]
    stencil = workload["stencil"]
    device = runtime.Device.get(0)
    library = runtime.load_library()
    tables = ()
    if hasattr(stencil, "intensity_lut"):
        tables = (stencil.distance_lut, stencil.intensity_lut)
    elif hasattr(stencil, "coefficients"):
        tables = (stencil.coefficients,)
    points = float(numpy.prod(shape))
    grid_buffer = device.alloc(int(points) * 4)
    out_buffer = device.alloc(int(points) * 4)
    utils = device.utils()
    runtime.launch(utils, "sc_fill_uniform", type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))(),
                   [ctypes.c_void_p(grid_buffer.ptr), ctypes.c_uint64(int(points)), ctypes.c_uint64(1),
                    ctypes.c_uint64(0), ctypes.c_float(workload["scale"])], device.stream)
    table_buffers = []
    for table in tables:
        table = numpy.ascontiguousarray(table)
        buffer = device.alloc(table.nbytes)
        runtime.check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(table.ctypes.data),
                                                  table.nbytes, device.stream))
        table_buffers.append(buffer)
    device.synchronize()

    class Fake(object):
        def __init__(self, shape, dtype):
            self.shape, self.dtype = shape, numpy.dtype(dtype)
    concrete = stencil.specializer.specialize((Fake(shape, numpy.float32),) + tuple(tables))
    pointers = [grid_buffer.ptr] + [b.ptr for b in table_buffers]
    for _ in range(3):
        concrete.launch(pointers, out_buffer.ptr)
    device.synchronize()
    start, stop = device.new_event(True), device.new_event(True)
    library.sc_event_record(start, device.stream)
    for _ in range(steps):
        concrete.launch(pointers, out_buffer.ptr)
    library.sc_event_record(stop, device.stream)
    device.synchronize()
    elapsed = ctypes.c_float()
    library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed))
    ms = elapsed.value / steps
    info = concrete.module.function_info(concrete.plan.function)
    geometry = concrete.plan.geometry(*concrete.plan.domain.sweep_range)
    print("{} {} {}: {:.4f} ms  {:.1f} GStencil/s  {:.0f} GB/s | {} grid={} regs={} env={}".format(
        name, boundary, "x".join(map(str, shape)), ms, points / ms / 1e6, 8 * points / ms / 1e6,
        concrete.plan.notes.get("config"), geometry.grid, info["regs"],
        {k: v for k, v in os.environ.items() if k.startswith("STENCIL_CUDA")}))


if __name__ == "__main__":
    main()
