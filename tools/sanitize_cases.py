#!/usr/bin/env python
"""A small tour through every generated kernel family, for compute-sanitizer (one tool per run):
direct, stream rank 3 (clamp/zero/copy, partial tiles), stream rank 2 (two blocks per thread),
tiled + re-rolled bilateral, fused two-sweep.  Each result is compared with the CPU oracle."""
import os

os.environ.setdefault("STENCIL_CUDA_TEMPORAL", "force")
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("STENCIL_CUDA_CACHE", "off")

from oracle.cref import reference_c                                        # noqa: E402
from stencil_code_b200.library.laplacian import LaplacianKernel            # noqa: E402
from stencil_code_b200.library.jacobi_3d import Jacobi3D                   # noqa: E402
from stencil_code_b200.library.simple import SimpleKernel                  # noqa: E402
from stencil_code_b200.library.convolution import ConvolutionFilter        # noqa: E402
from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter   # noqa: E402


def same(a, b):
    return bool((numpy.asarray(a).view(numpy.uint32) == b.view(numpy.uint32)).all())


def main():
    rng = numpy.random.RandomState(4)
    failures = 0
    os.environ["STENCIL_CUDA_STRATEGY"] = "stream"
    for mode in ("clamp", "zero", "copy"):
        grid = rng.random_sample((20, 37, 132)).astype(numpy.float32)
        ok = same(LaplacianKernel(boundary_handling=mode)(grid),
                  reference_c(LaplacianKernel(backend='python', boundary_handling=mode), grid))
        print("stream rank3 laplacian", mode, ok)
        failures += not ok
        grid = rng.random_sample((70, 1028)).astype(numpy.float32)
        ok = same(SimpleKernel(boundary_handling=mode)(grid),
                  reference_c(SimpleKernel(backend='python', boundary_handling=mode), grid))
        print("stream rank2 simple", mode, ok)
        failures += not ok
        kernel = numpy.ones((5, 5), numpy.int64)
        kernel[2, 2] = -4
        grid = rng.random_sample((40, 516)).astype(numpy.float32)
        ok = same(ConvolutionFilter(kernel, boundary_handling=mode)(grid),
                  reference_c(ConvolutionFilter(kernel, backend='python', boundary_handling=mode), grid))
        print("stream rank2 conv5", mode, ok)
        failures += not ok
        grid = rng.random_sample((14, 40, 260)).astype(numpy.float32)
        stencil = Jacobi3D(1.0, 0.25, boundary_handling=mode)
        python_side = Jacobi3D(1.0, 0.25, backend='python', boundary_handling=mode)
        expected = grid
        for _ in range(3):
            expected = reference_c(python_side, expected)
        ok = same(stencil.iterate(grid, 3), expected) and stencil.specializer.specialize((grid,)).temporal() is not None
        print("fused jacobi3d x3", mode, ok)
        failures += not ok
    grid = (rng.random_sample((40, 132)) * 255).astype(numpy.float32)
    ok = same(BetterBilateralFilter(3, 70)(grid), reference_c(BetterBilateralFilter(3, 70, backend='python'), grid))
    print("tiled bilateral r9", ok)
    failures += not ok
    os.environ["STENCIL_CUDA_STRATEGY"] = "direct"
    grid = rng.random_sample((9, 10, 11)).astype(numpy.float32)
    ok = same(LaplacianKernel()(grid), reference_c(LaplacianKernel(backend='python'), grid))
    print("direct laplacian", ok)
    failures += not ok
    print("SANITIZE_CASES", "PASSED" if failures == 0 else "FAILED")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
