"""bench.py leg for N > 1 GPUs: slab-decomposed iterated sweeps (strong scaling)."""
import ctypes
import os
import time

import numpy


def run(args, workload):
    import torch
    import torch.distributed as dist
    from . import runtime
    from .multi_gpu import SlabStencil
    import bench as bench_module

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world, rank = dist.get_world_size(), dist.get_rank()
    stencil = workload["stencil"]
    shape = workload["shape"]
    points = float(numpy.prod(shape))
    slab = SlabStencil(stencil, shape, device_index=local_rank)
    device, library = slab.device, slab.library
    slab.fill_uniform(seed=1, scale=workload["scale"])

    warmup = max(3, args.warmup)
    slab.iterate(warmup)
    slab.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    start, stop = device.new_event(True), device.new_event(True)
    with bench_module.ClockSampler(local_rank) as clocks:
        runtime.check(library.sc_event_record(start, device.stream))
        slab.iterate(args.steps)
        slab.join()
        runtime.check(library.sc_event_record(stop, device.stream))
        slab.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    elapsed = ctypes.c_float()
    runtime.check(library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed)))
    slowest = torch.tensor([elapsed.value], device="cuda")
    dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    ms_per_step = float(slowest.item()) / args.steps
    checksum = slab.checksum()

    # end to end, one sweep per step: pinned host slab -> device, sweep, result -> pinned host
    decomposition = slab.slab
    lo = max(0, decomposition.start - decomposition.first_real)
    hi = min(shape[0], decomposition.stop + decomposition.g)
    host_part = device.pinned_empty((hi - lo,) + tuple(shape[1:]), numpy.float32)
    host_part[...] = 0.5
    e2e_steps = 3

    class _Window(object):
        """lets SlabStencil.load slice [lo:hi] out of a host array that only holds this window"""
        def __init__(self, array, offset):
            self.array, self.offset = array, offset

        def __getitem__(self, planes):
            return self.array[planes.start - self.offset:planes.stop - self.offset]
    window = _Window(host_part, lo)
    for _ in range(2):                      # warm the pinned-buffer pool outside the timed region
        slab.load(window)
        slab.iterate(1)
        result = slab.result()
    dist.barrier()
    begin = time.perf_counter()
    for _ in range(e2e_steps):
        slab.load(window)
        slab.iterate(1)
        result = slab.result()
    dist.barrier()
    e2e_seconds = (time.perf_counter() - begin) / e2e_steps
    slowest_e2e = torch.tensor([e2e_seconds], device="cuda")
    dist.all_reduce(slowest_e2e, op=dist.ReduceOp.MAX)
    e2e_seconds = float(slowest_e2e.item())

    plan = slab.concrete.plan
    peak, peak_source = bench_module.hbm_peak()
    achieved = 8.0 * points / world / (ms_per_step * 1e-3) / 1e9          # per GPU
    line = None
    if rank == 0:
        line = {
            "metric": "GStencil/s", "value": points / (ms_per_step * 1e-3) / 1e9, "unit": "GStencil/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (counter-based uniform, generated on device)",
            "config": {"workload": workload["label"] + ", slab-decomposed along axis 0, {} planes per GPU, {}".format(
                           decomposition.planes,
                           "ghost planes written by the neighbour's kernel through NVLink peer memory"
                           if slab.exchange == "peer" else "ghost planes exchanged by NCCL send/recv (baseline mode)"),
                       "exchange": slab.exchange,
                       "sweeps_per_exchange": 2 if slab.fused is not None else 1,
                       "ghost_planes_per_side": decomposition.g,
                       "strategy": ("temporal2 (two sweeps fused per HBM pass and per ghost exchange)" if slab.fused is not None else plan.strategy), "l2": "per-GPU slab in+out larger than L2",
                       "kernel": plan.notes.get("config", {}), "launches_per_sweep": 4, "streams": "boundary planes + signal on an edge stream, interior on the main stream",
                       "checksum": "{:016x}".format(checksum)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_source,
                         "per": "GPU", "algorithmic_bytes_per_launch": 8.0 * points / world},
            "e2e": {"value": points / e2e_seconds / 1e9, "unit": "GStencil/s", "ms_per_step": e2e_seconds * 1e3,
                    "h2d_bytes_per_step": int(host_part.nbytes) * world,
                    "d2h_bytes_per_step": int(result.nbytes) * world,
                    "what": "per step: pinned host slab -> device, one sweep, slab -> pinned host, all ranks"},
            "gpu_launches": (args.steps // 2 + args.steps % 2 if slab.fused is not None else args.steps) * 4 * world,
            "clocks": clocks.summary(),
        }
    slab.close()
    dist.barrier()
    dist.destroy_process_group()
    return line
