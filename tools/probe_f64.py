#!/usr/bin/env python
"""float64 / 1-d throughput of the kernels that handle them: python tools/probe_f64.py"""
import ctypes, os, sys, time
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from cases import Stencil1d


def timed(stencil, tensor, label, bytes_per_point, iterate=0):
    for _ in range(3):
        out = stencil.iterate(tensor, iterate) if iterate else stencil(tensor)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    start.record()
    for _ in range(reps):
        out = stencil.iterate(tensor, iterate) if iterate else stencil(tensor)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps / max(1, iterate)
    points = tensor.numel()
    plan = stencil.specializer.specialize((tensor,)).plan
    print("{}: {} {} {:.3f} ms/sweep {:.1f} GStencil/s {:.0f} GB/s ({:.2f} of 6546)".format(
        label, plan.function, tuple(tensor.shape), ms, points / ms / 1e6, bytes_per_point * points / ms / 1e6,
        bytes_per_point * points / ms / 1e6 / 6545.9))


timed(DiagnosticStencil(), torch.rand((8192, 8192), device="cuda", dtype=torch.float64), "f64 rank 2 single", 16)
timed(DiagnosticStencil(), torch.rand((8192, 8192), device="cuda", dtype=torch.float64), "f64 rank 2 iterated x8", 16, iterate=8)
timed(DiagnosticStencil(), torch.rand((8192, 16384), device="cuda", dtype=torch.float32), "f32 rank 2 single", 8)
timed(Jacobi3D(1.0, 0.25), torch.rand((512, 512, 512), device="cuda", dtype=torch.float64), "f64 rank 3 single", 16)
timed(Stencil1d(), torch.rand((1 << 28,), device="cuda", dtype=torch.float32), "f32 rank 1 single", 8)
timed(Stencil1d(boundary_handling='copy'), torch.rand((1 << 28,), device="cuda", dtype=torch.float32), "f32 rank 1 copy", 8)
timed(Stencil1d(), torch.rand((1 << 27,), device="cuda", dtype=torch.float64), "f64 rank 1 single", 16)
