"""Run under torchrun on k GPUs: slab-decomposed sweeps must equal the single-GPU sweeps bit for
bit, and (small grid) the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py
"""
import os

os.environ.setdefault("STENCIL_CUDA_TEMPORAL", "force")     # the test planes are too small to be "worthwhile"
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from stencil_code_b200.backend.cuda.multi_gpu import SlabStencil
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
    from cases import COEFF27
    from oracle.cref import reference_c

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world, rank = dist.get_world_size(), dist.get_rank()
    failures = 0

    # 1. small grid against the CPU oracle, all boundary modes, uneven split (50 planes)
    for mode in ("clamp", "zero", "copy"):
        for make, tables, shape in (
                (lambda backend: Jacobi3D(1.0, 0.25, backend=backend, boundary_handling=mode), (), (50, 40, 128)),
                (lambda backend: SpecializedLaplacian27(backend=backend, boundary_handling=mode), (COEFF27,), (26, 24, 64))):
            grid = numpy.random.RandomState(5).random_sample(shape).astype(numpy.float32)
            sweeps = 7
            slab = SlabStencil(make('cuda'), shape, tables=tables, device_index=local_rank)
            slab.load(grid)
            slab.iterate(sweeps)
            slab.synchronize()
            mine = slab.result()
            parts = [None] * world
            dist.all_gather_object(parts, numpy.array(mine))
            total = slab.checksum()
            slab.close()
            if rank == 0:
                combined = numpy.concatenate(parts, axis=0)
                expected = grid
                python_side = make('python')
                for _ in range(sweeps):
                    expected = reference_c(python_side, expected, *tables)
                same = (combined.view(numpy.uint32) == expected.view(numpy.uint32)).all()
                single = SlabStencil(make('cuda'), shape, tables=tables, world=1, rank=0, device_index=local_rank,
                                     allow_fused=False)
                single.load(grid)
                single.iterate(sweeps)
                single.synchronize()
                same_sum = single.checksum() == total
                print("oracle check {} {} x{} ranks: {} checksum {}".format(
                    type(python_side).__name__, mode, world, "ok" if same else "MISMATCH",
                    "ok" if same_sum else "MISMATCH"))
                failures += (not same) + (not same_sum)
            dist.barrier()

    # 1b. a grid large enough for the fused two-sweep slab path (two ghost depths, one exchange
    #     per pair of sweeps) against the CPU oracle, odd sweep count
    shape, sweeps = (64, 128, 512), 5
    grid = numpy.random.RandomState(8).random_sample(shape).astype(numpy.float32)
    for mode in ("clamp", "copy"):
        slab = SlabStencil(Jacobi3D(1.0, 0.25, boundary_handling=mode), shape, device_index=local_rank)
        slab.load(grid)
        slab.iterate(sweeps)
        slab.synchronize()
        parts = [None] * world
        dist.all_gather_object(parts, numpy.array(slab.result()))
        layout = "fused, {} ghost planes".format(slab.slab.g) if slab.fused is not None else "single sweeps"
        slab.close()
        if rank == 0:
            expected = grid
            python_side = Jacobi3D(1.0, 0.25, backend='python', boundary_handling=mode)
            for _ in range(sweeps):
                expected = reference_c(python_side, expected, variant='c_parallel')
            same = (numpy.concatenate(parts, axis=0).view(numpy.uint32) == expected.view(numpy.uint32)).all()
            print("oracle check Jacobi3D {} 64x128x512 x{} sweeps, {} ranks ({}): {}".format(
                mode, sweeps, world, layout, "ok" if same else "MISMATCH"))
            failures += not same
        dist.barrier()

    # 1c. host -> ONE sweep -> host per rank (apply_host: the pipelined end-to-end path, ghost planes
    #     from the host window), then four more sweeps on the device (ghost planes refreshed first)
    shape = (96, 256, 512)
    grid = numpy.random.RandomState(9).random_sample(shape).astype(numpy.float32)
    slab = SlabStencil(Jacobi3D(1.0, 0.25), shape, device_index=local_rank)
    lo, hi = slab.host_window()
    first = numpy.array(slab.apply_host(numpy.ascontiguousarray(grid[lo:hi])))
    slab.iterate(4)
    parts_one, parts_five = [None] * world, [None] * world
    dist.all_gather_object(parts_one, first)
    dist.all_gather_object(parts_five, numpy.array(slab.result()))
    slab.close()
    if rank == 0:
        python_side = Jacobi3D(1.0, 0.25, backend='python')
        expected = reference_c(python_side, grid, variant='c_parallel')
        same_one = (numpy.concatenate(parts_one, axis=0).view(numpy.uint32) == expected.view(numpy.uint32)).all()
        for _ in range(4):
            expected = reference_c(python_side, expected, variant='c_parallel')
        same_five = (numpy.concatenate(parts_five, axis=0).view(numpy.uint32) == expected.view(numpy.uint32)).all()
        print("apply_host check 96x256x512, {} ranks: one sweep {}, then iterate(4) {}".format(
            world, "ok" if same_one else "MISMATCH", "ok" if same_five else "MISMATCH"))
        failures += (not same_one) + (not same_five)
    dist.barrier()

    # 2. larger grid (stream strategy, many chunks), k GPUs vs 1 GPU (plain single sweeps) by checksum
    shape = (256, 256, 512)
    slab = SlabStencil(Jacobi3D(1.0, 0.25), shape, device_index=local_rank)
    slab.fill_uniform(seed=3)
    slab.iterate(25)
    slab.synchronize()
    total = slab.checksum()
    strategy = "fused" if slab.fused is not None else slab.concrete.plan.strategy
    slab.close()
    if rank == 0:
        single = SlabStencil(Jacobi3D(1.0, 0.25), shape, world=1, rank=0, device_index=local_rank, allow_fused=False)
        single.fill_uniform(seed=3)
        single.iterate(25)
        single.synchronize()
        same = single.checksum() == total
        print("checksum check 256x256x512 x25 sweeps, {} ranks ({}): {}".format(world, strategy, "ok" if same else "MISMATCH"))
        failures += not same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASSED" if failures == 0 else "FAILED")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
