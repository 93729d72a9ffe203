#!/usr/bin/env python
"""Slab-decomposed iterated sweeps, timing only (run under torchrun):
   torchrun --nproc-per-node N tools/probe_slab.py [SWEEPS]   -> GStencil/s of C5 on N GPUs"""
import ctypes, os, sys
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from stencil_code_b200.backend.cuda import runtime
from stencil_code_b200.backend.cuda.multi_gpu import SlabStencil
from stencil_code_b200.library.jacobi_3d import Jacobi3D

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = (2048, 1024, 1024)
slab = SlabStencil(Jacobi3D(1.0, 0.25), shape, device_index=local)
slab.fill_uniform(seed=1)
slab.iterate(40)
slab.synchronize()
dist.barrier(); torch.cuda.synchronize()
device, library = slab.device, slab.library
best = None
for _ in range(3):
    start, stop = device.new_event(True), device.new_event(True)
    runtime.check(library.sc_event_record(start, device.stream))
    slab.iterate(sweeps)
    slab.join()
    runtime.check(library.sc_event_record(stop, device.stream))
    slab.synchronize()
    elapsed = ctypes.c_float()
    runtime.check(library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed)))
    cell = torch.tensor([elapsed.value], device="cuda")
    dist.all_reduce(cell, op=dist.ReduceOp.MAX)
    dist.barrier(); torch.cuda.synchronize()
    best = cell.item() if best is None else min(best, cell.item())
if dist.get_rank() == 0:
    plan = slab.fused[0] if slab.fused else slab.concrete.plan
    geometry = plan.step_geometry(slab.slab.first_real, slab.slab.first_real + slab.slab.planes) \
        if getattr(plan, "function_step", None) else None
    print("C5 on {} GPUs x{} sweeps: {:.4f} ms/sweep {:.1f} GStencil/s grid {} {}".format(
        dist.get_world_size(), sweeps, best / sweeps, float(numpy.prod(shape)) * sweeps / best / 1e6,
        geometry.grid if geometry else None, {k: v for k, v in os.environ.items() if k.startswith("STENCIL_CUDA")}))
slab.close()
dist.barrier()
dist.destroy_process_group()
