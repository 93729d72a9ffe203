"""bench.py leg for the iterated, slab-decomposed workload (C5 Jacobi-3D) at N = 1, 2, 4, 8 GPUs.

The same code path at every N: ``SlabStencil.iterate`` (fused two-sweep kernel where it applies);
at N = 1 the slab is the whole grid and there is no exchange.  Strong scaling: the grid is fixed,
slabs shrink with N."""
import ctypes
import os
import time

import numpy


def _timed_iterate(slab, steps, clock_sampler=None):
    """milliseconds of ``steps`` sweeps on this rank (CUDA events on the main stream, which joins
    the edge stream before the stop event)."""
    from . import runtime
    device, library = slab.device, slab.library
    start, stop = device.new_event(True), device.new_event(True)
    runtime.check(library.sc_event_record(start, device.stream))
    slab.iterate(steps)
    slab.join()
    runtime.check(library.sc_event_record(stop, device.stream))
    slab.synchronize()
    elapsed = ctypes.c_float()
    runtime.check(library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed)))
    return float(elapsed.value)


def one_gpu_reference(workload, warmup, steps, device_index):
    """The same workload, undecomposed, on one GPU: throughput and checksum after warmup + steps
    sweeps -- the denominator of a parallel efficiency and the parity reference of the
    decomposed run (same seed, same sweep count)."""
    from .multi_gpu import SlabStencil
    slab = SlabStencil(workload["stencil"], workload["shape"], world=1, rank=0, device_index=device_index)
    slab.fill_uniform(seed=1, scale=workload["scale"])
    slab.iterate(warmup)
    slab.synchronize()
    ms = _timed_iterate(slab, steps) / steps
    checksum = slab.checksum()
    points = float(numpy.prod(workload["shape"]))
    strategy = "temporal2" if slab.fused is not None else slab.concrete.plan.strategy
    slab.close()
    for buffer in slab.buffers:
        buffer.release()
    slab.device.trim()                       # hand the 16 GiB back before the decomposed run
    return {"value": points / ms / 1e6, "unit": "GStencil/s", "ms_per_step": ms, "strategy": strategy,
            "checksum": "{:016x}".format(checksum),
            "what": "this workload, undecomposed, on one GPU of this box, iterated, {} sweeps after {} of "
                    "warm-up and pre-roll, timed with CUDA events before the decomposed run".format(steps, warmup)}


def copy_floor(device, nbytes_in, nbytes_out, host_in, host_out):
    """ms of the raw transfers of one e2e step with nothing else going on: H2D of nbytes_in and
    D2H of nbytes_out running concurrently on two streams (the floor of a pipelined call)."""
    from . import runtime
    library = runtime.load_library()
    scratch_in, scratch_out = device.alloc(nbytes_in), device.alloc(nbytes_out)
    best = None
    for _ in range(2):
        device.synchronize()
        begin = time.perf_counter()
        runtime.check(library.sc_memcpy_h2d_async(ctypes.c_void_p(scratch_in.ptr), ctypes.c_void_p(host_in),
                                                  nbytes_in, device.h2d_stream))
        runtime.check(library.sc_memcpy_d2h_async(ctypes.c_void_p(host_out), ctypes.c_void_p(scratch_out.ptr),
                                                  nbytes_out, device.d2h_stream))
        device.synchronize(device.h2d_stream)
        device.synchronize(device.d2h_stream)
        seconds = time.perf_counter() - begin
        best = seconds if best is None else min(best, seconds)
    scratch_in.release()
    scratch_out.release()
    return best * 1e3


def run(args, workload):
    import bench as bench_module
    from . import runtime
    from .multi_gpu import SlabStencil

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        world = dist.get_world_size()
    rank = dist.get_rank() if dist else 0

    def barrier():
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def slowest(value):
        if not dist:
            return value
        cell = torch.tensor([value], device="cuda", dtype=torch.float64)
        dist.all_reduce(cell, op=dist.ReduceOp.MAX)
        return float(cell.item())

    stencil, shape = workload["stencil"], workload["shape"]
    points = float(numpy.prod(shape))
    warmup = max(3, args.warmup)
    steps = args.steps
    # untimed pre-roll of the same sweeps, about 0.2 s, inside the clock sampler's window: the timed
    # region (exactly `steps` sweeps) is only milliseconds long at 8 GPUs, too short for nvidia-smi
    preroll = 2 * int(round(0.1 / max(1e-6, 2.2e-3 * points / 2 ** 31 / world)))
    preroll = max(20, min(preroll, 2000))

    one_gpu = None
    if world > 1 and rank == 0 and not args.quick:
        try:
            one_gpu = one_gpu_reference(workload, warmup + preroll, steps, local_rank)
        except Exception as error:                         # must not lose the multi-GPU line
            one_gpu = {"error": str(error)[:200]}
    barrier()

    slab = SlabStencil(stencil, shape, device_index=local_rank)
    device = slab.device
    slab.fill_uniform(seed=1, scale=workload["scale"])
    slab.iterate(warmup)
    slab.synchronize()
    barrier()
    with bench_module.ClockSampler(local_rank) as clocks:
        slab.iterate(preroll)
        slab.synchronize()
        barrier()
        elapsed = _timed_iterate(slab, steps)
    barrier()
    ms_per_step = slowest(elapsed) / steps
    checksum = "{:016x}".format(slab.checksum())
    fused = slab.fused is not None
    decomposition = slab.slab

    # ---- end to end, ONE sweep per transfer: pinned host planes -> device, sweep, -> pinned host
    lo, hi = slab.host_window()
    host_in = device.pinned_empty((hi - lo,) + tuple(shape[1:]), numpy.float32)
    host_out = device.pinned_empty((decomposition.planes,) + tuple(shape[1:]), numpy.float32)
    # the current device grid (warm-up + timed sweeps applied) serves as host data: D2H of the window
    local_lo = lo - (decomposition.start - decomposition.first_real)
    library = slab.library
    runtime.check(library.sc_memcpy_d2h_async(
        ctypes.c_void_p(host_in.ctypes.data),
        ctypes.c_void_p(slab._current().ptr + local_lo * decomposition.plane_elems * 4), host_in.nbytes, device.stream))
    device.synchronize()
    e2e_steps = 3
    for _ in range(2):                                   # warm-up (first-touch of the pinned pages)
        slab.apply_host(host_in, out=host_out)
    floor_ms = slowest(copy_floor(device, host_in.nbytes, host_out.nbytes, host_in.ctypes.data, host_out.ctypes.data))
    barrier()
    begin = time.perf_counter()
    for _ in range(e2e_steps):
        slab.apply_host(host_in, out=host_out)
    e2e_seconds = slowest((time.perf_counter() - begin) / e2e_steps)
    barrier()

    # ---- end to end, the whole job of the BASELINE config: host -> 100 sweeps -> host
    job_sweeps = 100
    barrier()
    begin = time.perf_counter()
    slab.load(_Window(host_in, lo))
    slab.iterate(job_sweeps)
    slab.result(out=host_out)
    job_seconds = slowest(time.perf_counter() - begin)

    parity = None
    if world == 1 and not args.quick:
        # the fused two-sweep kernel against plain single sweeps, full grid, same seed and count
        if fused:
            plain = SlabStencil(stencil, shape, world=1, rank=0, device_index=local_rank, allow_fused=False)
            plain.fill_uniform(seed=1, scale=workload["scale"])
            plain.iterate(warmup + preroll + steps)
            parity = {"fused_vs_single_sweeps_checksum_equal": "{:016x}".format(plain.checksum()) == checksum}
            plain.close()
    elif one_gpu is not None and "checksum" in one_gpu:
        parity = {"parity_vs_one_gpu": one_gpu["checksum"] == checksum}

    plan = slab.fused[0] if fused else slab.concrete.plan
    peak, peak_source = bench_module.hbm_peak()
    sweeps_per_launch = 2 if fused else 1
    steps_with_exchange = steps // 2 + steps % 2 if fused else steps
    launches_per_step = slab.launches_per_step()
    bytes_per_launch = 8.0 * points / world * sweeps_per_launch          # the interior launch, per GPU
    launch_ms = ms_per_step * sweeps_per_launch
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = bench_module.measured_traffic("jacobi3d_iterated" if fused else "jacobi3d")
    if traffic is not None:
        traffic = traffic / world
    line = None
    if rank == 0:
        line = {
            "metric": "GStencil/s", "value": points / (ms_per_step * 1e-3) / 1e9, "unit": "GStencil/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (counter-based uniform [0,1), generated on device, seed 1)",
            "config": {"workload": workload["label"] + (
                           ", slab-decomposed along axis 0, {} planes per GPU".format(decomposition.planes)
                           if world > 1 else ", one GPU (the slab is the whole grid)"),
                       "step": "one sweep of the whole grid; pairs of sweeps are fused into one pass over HBM"
                               if fused else "one sweep of the whole grid",
                       "exchange": {"peer": "ghost planes written by the neighbour's kernel through NVLink peer memory",
                                    "nccl": "ghost planes exchanged by NCCL send/recv (baseline mode)",
                                    "none": "none (one GPU)"}[slab.exchange],
                       "sweeps_per_exchange": sweeps_per_launch,
                       "ghost_planes_per_side": decomposition.g if decomposition.ghosted else 0,
                       "strategy": "temporal2" if fused else plan.strategy,
                       "l2": "per-GPU slab in+out = {:.0f} MiB, larger than the 126 MiB L2 (no flush needed)".format(
                           2 * slab.nbytes / 2 ** 20),
                       "kernel": plan.notes.get("config", {}),
                       "launches_per_step_per_gpu": launches_per_step,
                       "preroll_sweeps": preroll,
                       "checksum": checksum},
            "roofline": {"bound": "hbm", "kernel": plan.function, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "dram_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "peak_source": peak_source, "per": "GPU",
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "sweeps_per_launch": sweeps_per_launch,
                         "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + "
                                           "dram__bytes_write.sum of one launch over the full grid, / n_gpus)",
                         "note": ("frac = algorithmic bytes (8 B/point/sweep x 2 sweeps per launch) / launch time / "
                                  "peak: above 1 because the fused kernel moves about half of them; dram_frac = "
                                  "bytes really moved (ncu) / launch time / peak") if fused else None},
            "e2e": {"value": points / e2e_seconds / 1e9, "unit": "GStencil/s", "ms_per_step": e2e_seconds * 1e3,
                    "h2d_bytes_per_step": int(host_in.nbytes) * world,
                    "d2h_bytes_per_step": int(host_out.nbytes) * world,
                    "sweeps_per_transfer": 1,
                    "raw_copy_floor_ms": floor_ms,
                    "what": "per step and rank: pinned host planes -> device, ONE sweep, real planes -> pinned host, "
                            "pipelined in chunks over three streams (SlabStencil.apply_host); max over ranks. "
                            "raw_copy_floor_ms = the same bytes moved by concurrent bare H2D + D2H copies",
                    "job_100_sweeps": {"value": points * job_sweeps / job_seconds / 1e9, "unit": "GStencil/s",
                                       "seconds": job_seconds, "sweeps_per_transfer": job_sweeps,
                                       "what": "the BASELINE job: host grid -> device, 100 sweeps, result -> host "
                                               "(load / iterate / result)"}},
            "gpu_launches": steps_with_exchange * launches_per_step * world,
            "clocks": clocks.summary(),
        }
        if parity:
            line["parity"] = parity
        if one_gpu is not None:
            line["one_gpu_same_workload"] = one_gpu
    del host_in, host_out
    slab.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return line


class _Window(object):
    """lets SlabStencil.load slice [lo:hi] out of a host array that only holds this window"""
    def __init__(self, array, offset):
        self.array, self.offset = array, offset

    def __getitem__(self, planes):
        return self.array[planes.start - self.offset:planes.stop - self.offset]
