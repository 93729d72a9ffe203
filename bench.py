#!/usr/bin/env python
"""bench.py -- GStencil/s of the stencil hot path on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl cuda|reference]

One JSON line on stdout (rank 0).  A "step" is one sweep of the workload's stencil over its grid.

Workloads (BASELINE.json configs):
  jacobi3d      C5  Jacobi3D 2048x1024x1024 float32, clamp, iterated; slab-decomposed over N GPUs
                    with halo exchange.  THE DEFAULT AT EVERY N (strong scaling: the grid is
                    fixed, slabs shrink; N = 1 runs the same code path on the whole grid)
  laplacian512  C2  LaplacianKernel 7-point, 512^3 float32, clamp, single pass
  simple1024    C1  SimpleKernel 3x3 Moore, 1024^2
  conv5         C4  ConvolutionFilter 5x5, 16384^2, --boundary zero|clamp|copy|wrap|periodic
  bilateral     C3  BetterBilateralFilter radius 9, 8192^2

Keys: `value` = device-resident throughput (CUDA events around K launches, inputs in HBM);
`e2e` = the same metric through the public API `stencil(numpy_array)` with pinned host buffers,
host->device and device->host copies inside the timed region; `roofline` = algorithmic bytes
(8 B/point) / kernel time against the measured HBM copy bandwidth; `cpu_baseline` = the restated
reference OpenMP backend (oracle/, kind "port") on this host's cores.
"""
from __future__ import print_function

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


# ---- workloads --------------------------------------------------------------------------------------
def make_workload(name, boundary, backend):
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.simple import SimpleKernel
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    if name == "laplacian512":
        return dict(label="C2 LaplacianKernel 7-point 512^3 float32 {} single pass".format(boundary or "clamp"),
                    stencil=LaplacianKernel(backend=backend, boundary_handling=boundary or "clamp"),
                    shape=(512, 512, 512), scale=1.0, cpu_variant="omp", cpu_shape=(512, 512, 512))
    if name == "jacobi3d":
        return dict(label="C5 Jacobi3D alpha=1 beta=0.25 2048x1024x1024 float32 {}".format(boundary or "clamp"),
                    stencil=Jacobi3D(1.0, 0.25, backend=backend, boundary_handling=boundary or "clamp"),
                    shape=(2048, 1024, 1024), scale=1.0, cpu_variant="omp", cpu_shape=(256, 1024, 1024))
    if name == "simple1024":
        return dict(label="C1 SimpleKernel 3x3 Moore 1024^2 float32 {}".format(boundary or "clamp"),
                    stencil=SimpleKernel(backend=backend, boundary_handling=boundary or "clamp"),
                    shape=(1024, 1024), scale=1000.0, cpu_variant="c_parallel", cpu_shape=(1024, 1024))
    if name == "conv5":
        kernel = numpy.ones((5, 5), numpy.int64)
        kernel[2, 2] = -4
        return dict(label="C4 ConvolutionFilter 5x5 16384^2 float32 {}".format(boundary or "copy"),
                    stencil=ConvolutionFilter(convolution_array=kernel, backend=backend,
                                              boundary_handling=boundary or "copy"),
                    shape=(16384, 16384), scale=1.0, cpu_variant="c_parallel", cpu_shape=(4096, 16384))
    if name == "bilateral":
        return dict(label="C3 BetterBilateralFilter sigma_d=3 (radius 9) sigma_i=70 8192^2 float32 clamp",
                    stencil=BetterBilateralFilter(3, 70, backend=backend),
                    shape=(8192, 8192), scale=255.0, cpu_variant="c_parallel", cpu_shape=(512, 8192))
    raise SystemExit("unknown workload " + name)


def host_grid(shape, scale, seed=1):
    rng = numpy.random.RandomState(seed)
    grid = rng.random_sample(shape).astype(numpy.float32)
    if scale != 1.0:
        grid *= numpy.float32(scale)
    return grid


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---- clocks ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.samples = []
        self.process = None
        self.device_index = device_index

    def __enter__(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read)
            self.thread.daemon = True
            self.thread.start()
            time.sleep(0.15)
        except Exception:
            self.process = None
        return self

    def _read(self):
        for line in self.process.stdout:
            self.samples.append(line.strip())

    def __exit__(self, *exc):
        if self.process is not None:
            time.sleep(0.1)
            self.process.terminate()
            try:
                self.process.wait(timeout=2)
            except Exception:
                self.process.kill()

    def summary(self):
        clocks, maxima, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clocks.append(float(parts[0]))
                maxima.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(c for c in clocks if c > 0.5 * max(clocks)) or sorted(clocks)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(maxima), "reasons": sorted(reasons),
                "samples": len(clocks)}


# ---- CPU legs --------------------------------------------------------------------------------------------
def host_threads():
    """Threads the CPU legs use: every core this process may run on."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def force_openmp_threads(count):
    """`torch.distributed.run` exports OMP_NUM_THREADS=1 to its workers: override it, in the
    environment (read when libgomp initialises) AND through the runtime call (if it already has).
    Returns omp_get_max_threads() as the OpenMP runtime the oracle links reports it."""
    os.environ["OMP_NUM_THREADS"] = str(count)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_dynamic(0)
        gomp.omp_set_num_threads(int(count))
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return None


def cpu_reference_run(workload, steps, warmup, min_seconds=0.0):
    """The restated reference OpenMP backend on this host's cores (oracle/, kind 'port').
    Runs `steps` sweeps, and keeps sweeping until `min_seconds` of CPU work have been timed."""
    cores = host_threads()
    force_openmp_threads(cores)
    from oracle.cref import ReferenceC, reference_args
    stencil = workload["stencil"]
    shape = workload["cpu_shape"]
    grid = host_grid(shape, workload["scale"])
    args = reference_args(stencil, (grid,))
    variant = workload["cpu_variant"]
    compiled = ReferenceC(stencil, args, variant)
    effective = force_openmp_threads(cores)          # the oracle's libgomp is loaded by now
    for _ in range(max(1, warmup)):
        compiled(*args)
    times = []
    while len(times) < steps or (sum(times) < min_seconds and len(times) < 2000):
        compiled(*args)
        times.append(compiled.last_seconds)
    seconds = sum(times) / len(times)
    points = float(numpy.prod(shape))
    return {"value": points / seconds / 1e9, "unit": "GStencil/s", "cores": effective or cores, "kind": "port",
            "variant": variant, "omp_get_max_threads": effective,
            "sample": "{} sweeps of {} = {:.1f} s of CPU work ({} backend restated from the reference, "
                      "gcc -O2 -std=c99 -fopenmp, omp_get_max_threads()={} on {} usable cores; the GPU arm's grid "
                      "is {})".format(len(times), "x".join(map(str, shape)), sum(times), variant, effective, cores,
                                      "x".join(map(str, workload["shape"]))),
            "ms_per_sweep": seconds * 1e3}


# ---- GPU legs ----------------------------------------------------------------------------------------------
class _Shape(object):
    """stands in for an array when only shape/dtype are needed to specialise a stencil"""
    def __init__(self, shape, dtype=numpy.float32):
        self.shape, self.dtype = tuple(shape), numpy.dtype(dtype)


def stencil_tables(stencil):
    if hasattr(stencil, "intensity_lut"):
        return (stencil.distance_lut, stencil.intensity_lut)
    if hasattr(stencil, "coefficients"):
        return (stencil.coefficients,)
    return ()


def device_sweeps(workload, steps, warmup, iterated=False):
    """Kernel-only timing on device-generated data: GStencil/s of `steps` sweeps (CUDA events).
    iterated=True ping-pongs the output back as input (the fused two-sweep kernel applies)."""
    from stencil_code_b200.backend.cuda import runtime
    stencil, shape = workload["stencil"], workload["shape"]
    device = runtime.Device.get(int(os.environ.get("LOCAL_RANK", "0")))
    library = runtime.load_library()
    points = int(numpy.prod(shape))
    tables = stencil_tables(stencil)
    concrete = stencil.specializer.specialize((_Shape(shape),) + tuple(tables))
    grid, out = device.alloc(points * 4), device.alloc(points * 4)
    fill = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
    runtime.launch(device.utils(), "sc_fill_uniform", fill,
                   [ctypes.c_void_p(grid.ptr), ctypes.c_uint64(points), ctypes.c_uint64(1), ctypes.c_uint64(0),
                    ctypes.c_float(workload["scale"])], device.stream)
    table_buffers = []
    for table in tables:
        table = numpy.ascontiguousarray(table)
        buffer = device.alloc(table.nbytes)
        runtime.check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(table.ctypes.data),
                                                  table.nbytes, device.stream))
        table_buffers.append(buffer)
    table_pointers = [b.ptr for b in table_buffers]

    def run(count):
        if iterated:
            concrete.sweeps(grid.ptr, out.ptr, table_pointers, count, device.stream)
        else:
            for _ in range(count):
                concrete.launch([grid.ptr] + table_pointers, out.ptr, stream=device.stream)
    run(max(3, warmup) + (max(3, warmup) % 2 if iterated else 0))
    device.synchronize()
    start, stop = device.new_event(True), device.new_event(True)
    runtime.check(library.sc_event_record(start, device.stream))
    run(steps)
    runtime.check(library.sc_event_record(stop, device.stream))
    device.synchronize()
    elapsed = ctypes.c_float()
    runtime.check(library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed)))
    for buffer in [grid, out] + table_buffers:
        buffer.release()
    ms = elapsed.value / steps
    fused = concrete.temporal() if iterated else None
    launches = steps if not fused else (steps // 2 + steps % 2)
    return {"gstencil_s": points / ms / 1e6, "ms_per_sweep": ms, "gb_s": 8.0 * points / ms / 1e6,
            "strategy": ("temporal2" if fused else concrete.plan.strategy), "launches": launches,
            "plan": concrete.plan}


def other_workloads(peak):
    """Kernel-only numbers of the remaining BASELINE configs (device-resident, synthetic data)."""
    rows = {}
    todo = [("C2 laplacian 512^3 clamp, single pass", "laplacian512", None, 100, False),
            ("C1 simple1024 clamp", "simple1024", None, 200, False),
            ("C3 bilateral r9 8192^2 clamp", "bilateral", None, 3, False),
            ("C4 conv5 16384^2 zero", "conv5", "zero", 10, False),
            ("C4 conv5 16384^2 clamp", "conv5", "clamp", 10, False),
            ("C4 conv5 16384^2 wrap (== zero, as in the reference)", "conv5", "wrap", 10, False),
            ("C4 conv5 16384^2 copy", "conv5", "copy", 10, False),
            ("C4 conv5 16384^2 periodic (true wrap-around: an opt-in addition, point kernel)", "conv5", "periodic", 4, False),
            ("C5 jacobi3d 2048x1024x1024 clamp, single sweeps (no temporal fusion)", "jacobi3d", None, 6, False)]
    for label, name, boundary, steps, iterated in todo:
        try:
            result = device_sweeps(make_workload(name, boundary, "cuda"), steps, 3, iterated)
            rows[label] = {"gstencil_s": round(result["gstencil_s"], 1), "ms_per_sweep": round(result["ms_per_sweep"], 4),
                           "frac_of_hbm_peak": round(result["gb_s"] / peak, 3), "strategy": result["strategy"]}
            if name == "laplacian512":
                rows[label]["e2e"] = laplacian_e2e()
            if name == "bilateral":
                # not HBM bound: 361 taps/point, each FSUB, FADD.RZ, LEA, LDS, FMUL, FMUL, FADD (7 instructions
                # the arithmetic needs; no FMA: the reference rounds each product) plus the row loads and
                # the loop: the executed count per tap is measured (ncu smsp__inst_executed, profiles/)
                # and set against the issue rate of 148 SMs x 4 schedulers x 32 lanes x 1.965 GHz
                per_tap = measured_traffic("bilateral_instr_per_tap") or 7.5
                taps = 361.0 * result["gstencil_s"] * 1e9
                peak_taps = 148 * 128 * 1.965e9 / per_tap
                rows[label]["instruction_roofline"] = {
                    "bound": "fp32 issue", "instr_per_tap": per_tap, "instr_per_tap_floor": 7,
                    "achieved_taps_per_s": taps, "peak_taps_per_s": peak_taps, "frac": round(taps / peak_taps, 3)}
        except Exception as error:       # a failing extra must not lose the headline line
            rows[label] = {"error": str(error)[:200]}
    return rows


def laplacian_e2e():
    """C2 through the public API: stencil(pinned numpy array) -> numpy array, copies inside."""
    from stencil_code_b200.backend.cuda import runtime
    workload = make_workload("laplacian512", None, "cuda")
    device = runtime.Device.get(int(os.environ.get("LOCAL_RANK", "0")))
    host_in = device.pinned_empty(workload["shape"], numpy.float32)
    host_in[...] = host_grid(workload["shape"], workload["scale"])
    stencil = workload["stencil"]
    for _ in range(2):
        stencil(host_in)
    begin = time.perf_counter()
    for _ in range(5):
        stencil(host_in)
    seconds = (time.perf_counter() - begin) / 5
    return {"value": float(numpy.prod(workload["shape"])) / seconds / 1e9, "unit": "GStencil/s",
            "ms_per_step": seconds * 1e3, "h2d_bytes_per_step": int(host_in.nbytes),
            "d2h_bytes_per_step": int(host_in.nbytes)}


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as handle:
            return json.load(handle).get(key)
    except Exception:
        return None


def run_single_gpu(args, workload):
    from stencil_code_b200.backend.cuda import runtime
    stencil = workload["stencil"]
    shape = workload["shape"]
    points = float(numpy.prod(shape))
    device = runtime.Device.get(int(os.environ.get("LOCAL_RANK", "0")))
    library = runtime.load_library()

    # inputs: pinned host grid (same seed as the parity tests use), resident copy on the device
    host_in = device.pinned_empty(shape, numpy.float32)
    host_in[...] = host_grid(shape, workload["scale"])
    call_args = (host_in,)
    if hasattr(stencil, "intensity_lut"):
        tables = (stencil.distance_lut, stencil.intensity_lut)
    elif hasattr(stencil, "coefficients"):
        tables = (stencil.coefficients,)
    else:
        tables = ()
    concrete = stencil.specializer.specialize(call_args + tables)
    plan = concrete.plan
    stream = device.stream
    buffers = [device.alloc(a.nbytes) for a in (host_in,) + tables]
    out_buffer = device.alloc(host_in.nbytes)
    for array, buffer in zip((host_in,) + tables, buffers):
        array = numpy.ascontiguousarray(array)
        runtime.check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(array.ctypes.data),
                                                  array.nbytes, stream))
    device.synchronize(stream)
    pointers = [b.ptr for b in buffers]

    def sweep():
        concrete.launch(pointers, out_buffer.ptr, stream=stream)

    for _ in range(max(3, args.warmup)):
        sweep()
    device.synchronize(stream)
    start, stop = device.new_event(True), device.new_event(True)
    with ClockSampler(device.index) as clocks:
        runtime.check(library.sc_event_record(start, stream))
        for _ in range(args.steps):
            sweep()
        runtime.check(library.sc_event_record(stop, stream))
        device.synchronize(stream)
    elapsed = ctypes.c_float()
    runtime.check(library.sc_event_elapsed_ms(start, stop, ctypes.byref(elapsed)))
    ms_per_step = elapsed.value / args.steps
    value = points / (ms_per_step * 1e-3) / 1e9

    # end to end through the public API: numpy (pinned) in, numpy out, copies inside the timed region
    e2e_steps = max(3, min(args.steps, 10))
    result = None
    for _ in range(2):
        result = stencil(host_in)
    begin = time.perf_counter()
    for _ in range(e2e_steps):
        result = stencil(host_in)
    e2e_seconds = (time.perf_counter() - begin) / e2e_steps
    checksum = float(result.ravel()[:: max(1, result.size // 4096)].sum())

    iterated = None
    peak, peak_source = hbm_peak()
    achieved = 8.0 * points / (ms_per_step * 1e-3) / 1e9
    info = concrete.module.function_info(plan.function)
    geometry = plan.geometry(*plan.domain.sweep_range)
    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (numpy RandomState(1) uniform)",
        "config": {"workload": workload["label"], "strategy": plan.strategy, "step": "one sweep of the grid",
                   "l2": "grid in+out = {:.0f} MiB, larger than the 126 MiB L2 (no flush needed)".format(
                       2 * host_in.nbytes / 2 ** 20) if 2 * host_in.nbytes > 2 * 126 * 2 ** 20 else
                   "grid fits in L2: L2-resident, launch-latency bound",
                   "kernel": dict(plan.notes.get("config", {}), grid=list(geometry.grid), block=list(geometry.block),
                                  regs=info["regs"], spill_bytes=info["local_bytes"])},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args.workload),
                     "dram_frac": (measured_traffic(args.workload) / (ms_per_step * 1e-3) / 1e9 / peak)
                     if measured_traffic(args.workload) else None,
                     "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + "
                                       "dram__bytes_write.sum per launch)",
                     "peak_source": peak_source, "algorithmic_bytes_per_launch": 8.0 * points * (2 if iterated and iterated["strategy"] == "temporal2" else 1),
                     "note": ("two sweeps per launch: algorithmic bytes are 8 B/point/sweep, the fused kernel "
                              "moves about half of that, so frac can exceed 1") if iterated else None},
        "e2e": {"value": points / e2e_seconds / 1e9, "unit": "GStencil/s", "ms_per_step": e2e_seconds * 1e3,
                "h2d_bytes_per_step": int(host_in.nbytes + sum(numpy.asarray(t).nbytes for t in tables)),
                "d2h_bytes_per_step": int(host_in.nbytes), "checksum": checksum},
        "gpu_launches": args.steps if not iterated else iterated["launches"],
        "clocks": clocks.summary(),
    }
    if iterated:
        line["config"]["iterated"] = {"strategy": iterated["strategy"], "single_sweep_gstencil_s": single_value}
    if not args.quick:
        line["workloads"] = other_workloads(peak)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_reference_run(make_workload(args.workload, args.boundary, "python"),
                                                 steps=3, warmup=1, min_seconds=10.0)
    return line


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=None)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--workload", default="jacobi3d")
    parser.add_argument("--boundary", default=None)
    parser.add_argument("--impl", default="cuda", choices=("cuda", "reference"))
    parser.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    parser.add_argument("--quick", action="store_true", help="skip the table of the other BASELINE configs")
    args = parser.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.steps is None:
        # BASELINE: 100 iterations of the Jacobi sweep; the others long enough for nvidia-smi to
        # sample clocks inside the timed region (sub-millisecond sweeps)
        args.steps = {"laplacian512": 400, "simple1024": 2000, "conv5": 200, "bilateral": 10,
                      "jacobi3d": 100}.get(args.workload, 20)

    if args.impl == "reference":
        if rank != 0:
            return
        force_openmp_threads(host_threads())             # before anything loads an OpenMP runtime
        workload = make_workload(args.workload, args.boundary, "python")
        base = cpu_reference_run(workload, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": "GStencil/s", "value": base["value"], "unit": "GStencil/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": base["ms_per_sweep"], "higher_is_better": True,
                "scaling": "strong" if args.workload == "jacobi3d" else "weak",   # as the cuda arm labels it
                "vs_baseline": None, "dtype": "f32", "data": "synthetic (numpy RandomState(1) uniform)",
                "config": {"workload": workload["label"], "cpu_grid": list(workload["cpu_shape"]),
                           "step": "one sweep of the CPU sample grid; GStencil/s is per point, so it compares "
                                   "with the GPU arm's full grid"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "GStencil/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    workload = make_workload(args.workload, args.boundary, "cuda")
    if args.workload == "jacobi3d":
        from stencil_code_b200.backend.cuda import multi_gpu_bench
        line = multi_gpu_bench.run(args, workload)
        if line is None:
            return
        if line["n_gpus"] == 1:
            peak, _ = hbm_peak()
            if not args.quick:
                line["workloads"] = other_workloads(peak)
            if not args.no_cpu:
                line["cpu_baseline"] = cpu_reference_run(make_workload(args.workload, args.boundary, "python"),
                                                         steps=3, warmup=1, min_seconds=10.0)
        print(json.dumps(line))
        return
    if args.gpus > 1 or world > 1:
        raise SystemExit("only the iterated workload (jacobi3d) is decomposed over several GPUs; single-pass "
                         "configs have no exchange step (replicas only)")
    print(json.dumps(run_single_gpu(args, workload)))


if __name__ == "__main__":
    main()
