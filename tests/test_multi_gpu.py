"""Slab decomposition: host logic on CPU with gloo (world_size 2 and 3), device path on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from stencil_code_b200.backend.cuda.multi_gpu import SlabDecomposition
from stencil_code_b200.stencil_exception import StencilException

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decomposition_arithmetic():
    parts = [SlabDecomposition((50, 8, 8), 1, 4, r) for r in range(4)]
    assert [p.planes for p in parts] == [13, 13, 12, 12]
    assert [p.start for p in parts] == [0, 13, 26, 38] and parts[-1].stop == 50
    assert parts[0].local_shape == (15, 8, 8) and not parts[0].open_lo and parts[0].open_hi
    top, interior, bottom = parts[1].ranges()
    assert top == (1, 2) and interior == (2, 13) and bottom == (13, 14)
    top, interior, bottom = parts[0].ranges()
    assert top == (1, 1) and interior == (1, 13) and bottom == (13, 14)     # global top: no exchange
    assert parts[1].sends() == [(0, (1, 2), (14, 15)), (2, (13, 14), (0, 1))]
    single = SlabDecomposition((50, 8, 8), 1, 1, 0)
    assert single.local_shape == (50, 8, 8) and single.ranges() == ((0, 0), (0, 50), (50, 50))
    with pytest.raises(StencilException):
        SlabDecomposition((6, 8, 8), 2, 4, 0)
    # two ghost depths per side: the layout of the fused two-sweep path
    deep = SlabDecomposition((50, 8, 8), 1, 4, 1, ghost_planes=2)
    assert deep.local_shape == (17, 8, 8) and deep.first_real == 2
    assert deep.ranges() == ((2, 4), (4, 13), (13, 15))
    assert deep.sends() == [(0, (2, 4), (15, 17)), (2, (13, 15), (0, 2))]


WORKER = r'''
import os, sys
import numpy
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
from stencil_code_b200.backend.cuda.multi_gpu import SlabDecomposition
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from oracle.cref import reference_c

dist.init_process_group("gloo")
world, rank = dist.get_world_size(), dist.get_rank()
shape, sweeps = (23, 10, 12), 5
stencil = Jacobi3D(1.0, 0.25, backend='python', boundary_handling='clamp')
grid = numpy.random.RandomState(2).random_sample(shape).astype(numpy.float32)
slab = SlabDecomposition(shape, stencil.ghost_depth[0], world, rank)
h = slab.h
local = numpy.zeros(slab.local_shape, numpy.float32)
lo, hi = max(0, slab.start - h), min(shape[0], slab.stop + h)
local[lo - (slab.start - h):hi - (slab.start - h)] = grid[lo:hi]
for _ in range(sweeps):
    # a rank computes on [first real plane or its open-side ghosts .. ]: clamping at an open face
    # only corrupts the outermost ghost plane, which is discarded
    first = 0 if slab.open_lo else slab.first_real
    last = local.shape[0] if slab.open_hi else slab.first_real + slab.planes
    swept = reference_c(stencil, numpy.ascontiguousarray(local[first:last]))
    new = numpy.zeros_like(local)
    new[first:last] = swept
    requests = []
    for neighbour, source, destination in slab.sends():
        requests.append(dist.isend(torch.from_numpy(numpy.ascontiguousarray(new[source[0]:source[1]])), neighbour))
    for neighbour, source, destination in slab.sends():
        # what I receive from `neighbour` lands where the neighbour's plan says; by symmetry that is
        # my ghost range on that side
        ghost = (0, h) if neighbour < rank else (h + slab.planes, 2 * h + slab.planes)
        buffer = torch.empty((h,) + tuple(shape[1:]), dtype=torch.float32)
        dist.recv(buffer, neighbour)
        new[ghost[0]:ghost[1]] = buffer.numpy()
    for request in requests:
        request.wait()
    local = new
parts = [None] * world
dist.all_gather_object(parts, numpy.array(local[slab.first_real:slab.first_real + slab.planes]))
if rank == 0:
    expected = grid
    for _ in range(sweeps):
        expected = reference_c(stencil, expected)
    combined = numpy.concatenate(parts, axis=0)
    assert (combined.view(numpy.uint32) == expected.view(numpy.uint32)).all(), "slab result differs"
    print("SLAB_OK", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_slab_logic_with_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    done = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
         "--master-addr", "127.0.0.1", "--master-port", str(29530 + world), str(script)],
        capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert done.returncode == 0, done.stdout[-2000:] + done.stderr[-3000:]
    assert "SLAB_OK {}".format(world) in done.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_slabs_on_two_gpus_match_single_gpu_and_oracle(exchange):
    """ghost planes stored by the neighbour's kernel through peer memory (the product's way), and
    the NCCL send/recv baseline mode: both bit-identical to one GPU and to the oracle"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    done = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29541" if exchange == "peer" else "29543",
         os.path.join(ROOT, "tests", "multi_gpu_check.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, STENCIL_CUDA_EXCHANGE=exchange))
    assert done.returncode == 0 and "MULTI_GPU_CHECK PASSED" in done.stdout, \
        done.stdout[-3000:] + done.stderr[-3000:]


def test_step_kernel_geometry_covers_the_slab():
    """One launch per exchange step: the chunk at each open end of the slab is a z-slice of its
    own (scheduled first, holding every plane the neighbour receives), the middle chunks follow;
    together they tile the real planes exactly once -- thick and thin slabs, one or two neighbours."""
    import numpy
    from stencil_code_b200.backend.cuda import codegen_temporal, planner, runtime
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.stencil_kernel import StencilArgConfig
    specializer = Jacobi3D(1.0, 0.25).specializer
    ghost = 2
    for planes, (n1, n2) in ((4, (256, 512)), (5, (256, 512)), (12, (256, 512)), (32, (256, 512)), (70, (256, 512)),
                             (130, (1024, 1024)), (256, (1024, 1024))):
        for faces in ((True, True), (False, True), (True, False)):
            shape = (planes + 2 * ghost, n1, n2)
            plan = specializer.transform((StencilArgConfig(shape[0], numpy.dtype('float32'), 3, shape),),
                                         open_faces=faces, slab_ghost=ghost)
            fused = codegen_temporal.make_plan(plan.program, plan.domain, planner.B200, plan.options)
            assert fused.function_step == "stencil_temporal2_step"
            launch = fused.step_geometry(ghost, ghost + planes)
            top, middle, bottom = launch.ranges
            assert top[0] == ghost and bottom[1] == ghost + planes and top[1] == middle[0] and middle[1] == bottom[0]
            assert (top[1] - top[0] >= ghost) if faces[0] else top[1] == top[0]
            assert (bottom[1] - bottom[0] >= ghost) if faces[1] else bottom[1] == bottom[0]
            middle_chunks = -(-(middle[1] - middle[0]) // launch.extra[0]) if middle[1] > middle[0] else 0
            assert launch.grid[2] == sum(faces) + middle_chunks
    # the three entry points cross-compile for sm_100a (with the griddepcontrol instructions in)
    module = runtime.compile_source(fused.source, "step_check.cu", list(fused.options) + ["-DSC_PDL"])
    assert module.cubin()[:4] == b"\x7fELF"


def test_host_windows_cover_what_a_sweep_reads():
    """apply_host: the planes a rank takes from the host are its real planes plus the ghost planes
    towards each neighbour, clipped to the grid; neighbouring windows overlap by 2 g planes."""
    for world, ghost in ((1, None), (2, None), (4, 2), (8, 2)):
        parts = [SlabDecomposition((2048, 8, 8), 1, world, r, ghost_planes=ghost) for r in range(world)]
        for r, part in enumerate(parts):
            lo, hi = part.host_window()
            g = part.g if world > 1 else 0
            assert lo == max(0, part.start - g) and hi == min(2048, part.stop + g)
            assert hi - lo <= part.local_shape[0]
            local_lo = lo - (part.start - part.first_real)          # where the window starts in the local array
            assert local_lo == (0 if r > 0 or world == 1 else part.first_real)
            if r + 1 < world:
                assert hi - parts[r + 1].host_window()[0] == 2 * g


def test_pipelined_host_sweep_copies_and_sweeps_every_plane_once():
    """ConcreteCudaStencil.sweep_host with a recording stand-in for the C ABI: whatever the chunking,
    every input plane of the window is copied to the device exactly once, the launches tile the
    sweep range in order, every output plane is copied back exactly once, and no chunk is swept
    before the planes it reads (its own and the first ghost planes of the next chunk) were enqueued."""
    import numpy
    from stencil_code_b200.backend.cuda import runtime

    class Recorder(object):
        def __init__(self):
            self.log = []

        def __getattr__(self, name):
            def call(*args):
                self.log.append((name,) + tuple(a.value if hasattr(a, "value") else a for a in args))
                return 0
            return call

    class FakeDevice(object):
        stream, h2d_stream, d2h_stream = "compute", "h2d", "d2h"

        def __init__(self):
            self.events = 0

        def event(self):
            self.events += 1
            return "event{}".format(self.events)

        def recycle_event(self, event):
            pass

        def synchronize(self, stream=None):
            pass

    class Program(object):
        shape, out_dtype_size, ghost_depth = (300, 512, 512), 4, (2, 1, 1)

    recorder = Recorder()
    original = runtime.load_library
    runtime.load_library = lambda: recorder
    try:
        stencil = object.__new__(runtime.ConcreteCudaStencil)
        stencil.device, stencil.program = FakeDevice(), Program()
        launches = []
        stencil.launch = lambda pointers, out, begin, end, stream=None: (
            launches.append((begin, end)), recorder.log.append(("launch", begin, end)))
        plane = 512 * 512 * 4
        # a slab window: local planes 0..300 come from the host, planes 2..298 are swept
        stencil.sweep_host(1 << 40, (0, 300), [1 << 30], 1 << 31, (1 << 41) - 2 * plane, (2, 298))
    finally:
        runtime.load_library = original
    copied_in = numpy.zeros(300, int)
    copied_out = numpy.zeros(300, int)
    landed = numpy.zeros(300, bool)
    for entry in recorder.log:
        if entry[0] == "sc_memcpy_h2d_async":
            first = (entry[1] - (1 << 30)) // plane
            count = entry[3] // plane
            assert (entry[2] - (1 << 40)) // plane == first            # host offset matches the device offset
            copied_in[first:first + count] += 1
            landed[first:first + count] = True
        elif entry[0] == "launch":
            begin, end = entry[1], entry[2]
            assert landed[max(0, begin - 2):min(300, end + 2)].all(), "swept before its input was enqueued"
        elif entry[0] == "sc_memcpy_d2h_async":
            first = (entry[2] - (1 << 31)) // plane
            assert (entry[1] - ((1 << 41) - 2 * plane)) // plane == first
            copied_out[first:first + entry[3] // plane] += 1
    assert (copied_in == 1).all()
    assert (copied_out[2:298] == 1).all() and copied_out[:2].sum() == 0 and copied_out[298:].sum() == 0
    assert launches[0][0] == 2 and launches[-1][1] == 298 and len(launches) > 4
    assert all(a[1] == b[0] for a, b in zip(launches, launches[1:]))
