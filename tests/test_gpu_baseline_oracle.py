"""Oracle parity at BASELINE.json's full sizes: backend='cuda' against the restated reference C
(``oracle/``, C-backend boundary semantics, outer loop parallel) on the SAME inputs, bit patterns,
full grids (halo included).  Tolerance: none.

  C1  SimpleKernel 1024^2 (Moore and the README's literal list), clamp
  C2  LaplacianKernel 512^3, clamp and copy                    (benchmarks/laplacian.py:12-20)
  C3  BetterBilateralFilter r=9 8192^2: the GPU filters the whole image; two 1024-row bands (top
      edge, middle) are checked against the oracle run on the band plus its 9-row aprons
  C4  ConvolutionFilter 5x5 16384^2 x zero / clamp / wrap / copy, and the opt-in 'periodic' mode
      against numpy.pad(mode='wrap') + the oracle               (benchmarks/convolve.py:16-38)
  C5  Jacobi-3D: 256x1024x1024 one sweep and iterate(2) (fused kernel); and the FULL
      2048x1024x1024 grid, two fused sweeps, checked slab by slab: eight 256-plane slabs, the
      oracle applied twice to each slab plus two planes of apron per side (the apron absorbs the
      oracle's clamp at the cut; at the global faces the cut IS the clamp)

Every call goes through the public API (numpy in / numpy out), i.e. through the C ABI.
"""
import os

import numpy
import pytest

from cases import oracle_output, periodic_oracle

pytestmark = pytest.mark.gpu


def random_grid(shape, seed=1, scale=1.0):
    grid = numpy.random.default_rng(seed).random(shape, dtype=numpy.float32)
    if scale != 1.0:
        grid *= numpy.float32(scale)
    return grid


def assert_same_bits(actual, expected, what=""):
    actual, expected = numpy.ascontiguousarray(actual), numpy.ascontiguousarray(expected)
    assert actual.shape == expected.shape and actual.dtype == expected.dtype == numpy.float32
    different = actual.view(numpy.uint32) != expected.view(numpy.uint32)
    if different.any():
        first = tuple(numpy.argwhere(different)[0])
        raise AssertionError("{}: {} of {} elements differ; first at {}: got {!r}, expected {!r}".format(
            what, int(different.sum()), different.size, first, actual[first], expected[first]))


def available_host_gib():
    try:
        with open("/proc/meminfo") as handle:
            for line in handle:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


@pytest.mark.parametrize("kernel", ["moore", "readme"])
def test_c1_simple_kernel_1024(kernel):
    from stencil_code_b200.library.simple import SimpleKernel, ReadmeSimpleKernel
    cls = SimpleKernel if kernel == "moore" else ReadmeSimpleKernel
    grid = random_grid((1024, 1024), scale=1000.0)
    expected = oracle_output(cls(backend='python'), grid, variant='c_parallel')
    assert_same_bits(cls(backend='cuda')(grid), expected, "C1 " + kernel)


@pytest.mark.parametrize("mode", ["clamp", "copy"])
def test_c2_laplacian_512(mode):
    from stencil_code_b200.library.laplacian import LaplacianKernel
    grid = random_grid((512, 512, 512))
    stencil = LaplacianKernel(backend='cuda', boundary_handling=mode)
    assert stencil.specializer.specialize((grid,)).plan.strategy == 'stream'
    expected = oracle_output(LaplacianKernel(backend='python', boundary_handling=mode), grid, variant='c_parallel')
    assert_same_bits(stencil(grid), expected, "C2 " + mode)


def test_c3_bilateral_8192_bands():
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    grid = random_grid((8192, 8192), scale=255.0)
    result = BetterBilateralFilter(3, 70, backend='cuda')(grid)
    python_side = BetterBilateralFilter(3, 70, backend='python')
    radius, band = 9, 1024
    # top band: the grid's first rows, clamp binds at row 0 exactly as in the full image
    expected = oracle_output(python_side, numpy.ascontiguousarray(grid[:band + radius]), variant='c_parallel')
    assert_same_bits(result[:band], expected[:band], "C3 rows 0..1023")
    # a middle band with aprons on both sides
    first = 4000
    expected = oracle_output(python_side, numpy.ascontiguousarray(grid[first - radius:first + band + radius]),
                             variant='c_parallel')
    assert_same_bits(result[first:first + band], expected[radius:radius + band], "C3 rows 4000..5023")
    # bottom edge
    expected = oracle_output(python_side, numpy.ascontiguousarray(grid[-(256 + radius):]), variant='c_parallel')
    assert_same_bits(result[-256:], expected[-256:], "C3 last 256 rows")


def _conv5(mode, backend):
    from stencil_code_b200.library.convolution import ConvolutionFilter
    kernel = numpy.ones((5, 5), numpy.int64)
    kernel[2, 2] = -4                                                     # benchmarks/convolve.py:16-24
    return ConvolutionFilter(convolution_array=kernel, backend=backend, boundary_handling=mode)


@pytest.mark.parametrize("mode", ["zero", "clamp", "wrap", "copy"])
def test_c4_convolution_16384(mode):
    grid = random_grid((16384, 16384))
    stencil = _conv5(mode, 'cuda')
    expected = oracle_output(_conv5(mode, 'python'), grid, variant='c_parallel')
    assert_same_bits(stencil(grid), expected, "C4 " + mode)


def test_c4_convolution_16384_periodic():
    grid = random_grid((16384, 16384))
    expected = periodic_oracle(lambda mode: _conv5(mode, 'python'), grid)
    assert_same_bits(_conv5('periodic', 'cuda')(grid), expected, "C4 periodic")


@pytest.mark.parametrize("mode", ["clamp", "copy"])
def test_c5_jacobi_slab_one_and_two_sweeps(mode):
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    grid = random_grid((256, 1024, 1024))
    stencil = Jacobi3D(1.0, 0.25, backend='cuda', boundary_handling=mode)
    python_side = Jacobi3D(1.0, 0.25, backend='python', boundary_handling=mode)
    once = oracle_output(python_side, grid, variant='c_parallel')
    assert_same_bits(stencil(grid), once, "C5 one sweep " + mode)
    twice = oracle_output(python_side, once, variant='c_parallel')
    assert stencil.specializer.specialize((grid,)).temporal() is not None, "fused kernel expected at this size"
    assert_same_bits(stencil.iterate(grid, 2), twice, "C5 iterate(2) " + mode)
    assert_same_bits(stencil.iterate(grid, 3), oracle_output(python_side, twice, variant='c_parallel'),
                     "C5 iterate(3) " + mode)


def test_c5_jacobi_full_grid_against_oracle_slab_by_slab():
    """The whole 2^31-point grid: two sweeps by the fused kernel, every plane compared with the
    oracle.  The oracle works on 256-plane slabs with a 2-plane apron per side (two sweeps reach two
    planes); apron planes, where its clamp at the cut differs from the real neighbour, are dropped."""
    if available_host_gib() < 40:
        pytest.skip("needs about 20 GiB of host memory for the 8 GiB grid and its result")
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    shape = (2048, 1024, 1024)
    grid = numpy.empty(shape, numpy.float32)
    for first in range(0, shape[0], 256):                      # generated in pieces: no 16 GiB temporaries
        grid[first:first + 256] = numpy.random.default_rng(100 + first).random((256,) + shape[1:], dtype=numpy.float32)
    stencil = Jacobi3D(1.0, 0.25, backend='cuda')
    result = stencil.iterate(grid, 2)
    assert stencil.specializer.specialize((grid,)).temporal() is not None
    python_side = Jacobi3D(1.0, 0.25, backend='python')
    apron = 2
    for first in range(0, shape[0], 256):
        lo, hi = max(0, first - apron), min(shape[0], first + 256 + apron)
        piece = numpy.ascontiguousarray(grid[lo:hi])
        expected = oracle_output(python_side, oracle_output(python_side, piece, variant='c_parallel'),
                                 variant='c_parallel')
        assert_same_bits(result[first:first + 256], expected[first - lo:first - lo + 256],
                         "C5 full grid, planes {}..{}".format(first, first + 255))
    # and single sweeps (the streaming kernel) on the same grid: first and last slab
    once = stencil(grid)
    for first in (0, shape[0] - 256):
        lo, hi = max(0, first - 1), min(shape[0], first + 256 + 1)
        expected = oracle_output(python_side, numpy.ascontiguousarray(grid[lo:hi]), variant='c_parallel')
        assert_same_bits(once[first:first + 256], expected[first - lo:first - lo + 256],
                         "C5 full grid single sweep, planes {}..{}".format(first, first + 255))
