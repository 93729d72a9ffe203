"""The stencil cases shared by the oracle, code-generation and GPU parity tests.

Each case mirrors one used by the reference's own tests (``test/test_c_end_to_end.py``,
``test/test_boundary_handling.py``) or its README, built from THIS repo's library classes.
``make(boundary, backend)`` returns the stencil; ``extra`` are the arguments after the grid.
"""
import numpy

from stencil_code_b200.neighborhood import Neighborhood
from stencil_code_b200.stencil_kernel import Stencil
from stencil_code_b200.library.laplacian import LaplacianKernel
from stencil_code_b200.library.jacobi_stencil import Jacobi
from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
from stencil_code_b200.library.two_d_heat import TwoDHeatFlow
from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
from stencil_code_b200.library.convolution import ConvolutionFilter
from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
from stencil_code_b200.library.bilateral_filter import BilateralFilter, gaussian
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.simple import SimpleKernel, ReadmeSimpleKernel

ALL_MODES = ('clamp', 'zero', 'copy', 'wrap')


class Stencil1d(Stencil):                      # reference test/test_boundary_handling.py:145-151
    neighborhoods = [[(-1,), (0,), (1,)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            for y in self.neighbors(x, 0):
                out_grid[x] += 2 * in_grid[y]


class Clamper(Stencil):                        # reference test/test_boundary_handling.py:15-21
    neighborhoods = [Neighborhood.von_neuman_neighborhood(radius=1, dim=2)]

    def kernel(self, in_grid, out_grid):
        for p in self.interior_points(out_grid):
            for n in self.neighbors(p, 0):
                out_grid[p] += in_grid[n]


class OddWeights(Stencil):
    """Double weights that are NOT powers of two: must stay in double to match C bit for bit."""
    neighborhoods = [Neighborhood.moore_neighborhood(radius=1, dim=2)]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = 0.3 * in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += 0.1 * in_grid[y]


class Radius2Cross3D(Stencil):
    """Radius-2 von Neumann cross in 3-d with a weight table and a local temporary."""
    neighborhoods = [Neighborhood.von_neuman_neighborhood(radius=2, dim=3)]

    def distance(self, x, y):
        return sum(abs(x[i] - y[i]) for i in range(len(x)))

    def kernel(self, in_grid, weights, out_grid):
        for x in self.interior_points(out_grid):
            for y in self.neighbors(x, 0):
                out_grid[x] += weights[self.distance(x, y)] * in_grid[y]


class WideStar3D(Stencil):
    """Star in z with in-plane reach 2 (ghost depth (1, 2, 2)): fused with a radius-1 stencil, the
    two stages have different halos."""
    neighborhoods = [[(1, 0, 0), (-1, 0, 0)], [(0, -2, 0), (0, 2, 0), (0, 0, -2), (0, 0, 2), (0, 1, 1)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = 0.5 * in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += 0.125 * in_grid[y]
            for y in self.neighbors(x, 1):
                out_grid[x] -= 0.0625 * in_grid[y]


CONV5 = numpy.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, -4, 1, 3],
                     [1, 1, 1, 1, 1], [1, 1, 1, 7, 1]])        # test/test_c_end_to_end.py:49-57
CONV5_BENCH = numpy.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, -4, 1, 1],
                           [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]])  # benchmarks/convolve.py:16-24
COEFF27 = numpy.array([1.0, 0.5, 0.25, 0.125]).astype(numpy.float32)
WEIGHTS3 = numpy.array([0.5, 0.25, 0.125]).astype(numpy.float32)


def _plain(cls):
    return lambda boundary, backend: cls(backend=backend, boundary_handling=boundary)


class Case(object):
    def __init__(self, name, make, extra=(), shape=(8, 8), scale=1000, modes=ALL_MODES, golden=True):
        self.name, self.make, self.extra = name, make, tuple(extra)
        self.shape, self.scale, self.modes, self.golden = tuple(shape), scale, modes, golden

    def grid(self, seed=0, shape=None, dtype=numpy.float32):
        rng = numpy.random.RandomState(seed + sum(ord(c) for c in self.name))
        return (rng.random_sample(shape or self.shape) * self.scale).astype(dtype)


CASES = [
    Case("laplacian", _plain(LaplacianKernel), shape=(9, 10, 12)),
    Case("jacobi2d", _plain(Jacobi), shape=(17, 20)),
    Case("diagnostic", _plain(DiagnosticStencil), shape=(16, 12)),
    Case("two_d_heat", _plain(TwoDHeatFlow), shape=(6, 9, 8)),
    Case("laplacian27", _plain(SpecializedLaplacian27), extra=(COEFF27,), shape=(8, 9, 12), scale=255),
    Case("stencil1d", _plain(Stencil1d), shape=(24,)),
    Case("clamper", _plain(Clamper), shape=(10, 10)),
    Case("readme_simple", _plain(ReadmeSimpleKernel), shape=(12, 16)),
    Case("convolution5", lambda b, backend: ConvolutionFilter(convolution_array=CONV5, backend=backend),
         shape=(20, 24), modes=('copy',)),
    Case("better_bilateral", lambda b, backend: BetterBilateralFilter(sigma_d=1, sigma_i=70, backend=backend),
         shape=(20, 16), scale=255, modes=('clamp',)),
    Case("bilateral_r3", lambda b, backend: BilateralFilter(radius=3, backend=backend),
         extra=(gaussian(1, 6), gaussian(70, 256)), shape=(12, 16), scale=255, modes=('clamp',)),
    # cases without a reference-made fixture (the reference has no such class); oracle-only
    Case("simple_moore", _plain(SimpleKernel), shape=(12, 16), golden=False),
    Case("jacobi3d", lambda b, backend: Jacobi3D(1.0, 0.25, backend=backend, boundary_handling=b),
         shape=(10, 8, 12), scale=1, golden=False),
    Case("odd_weights", _plain(OddWeights), shape=(14, 12), golden=False),
    Case("cross3d_r2", _plain(Radius2Cross3D), extra=(WEIGHTS3,), shape=(9, 10, 12), golden=False),
    Case("bilateral_r9", lambda b, backend: BetterBilateralFilter(sigma_d=3, sigma_i=70, backend=backend),
         shape=(40, 128), scale=255, modes=('clamp',), golden=False),
    Case("convolution5_modes",
         lambda b, backend: ConvolutionFilter(convolution_array=CONV5_BENCH, backend=backend, boundary_handling=b),
         shape=(20, 24), golden=False),
]

CASE_MODE_PAIRS = [(case, mode) for case in CASES for mode in case.modes]
PAIR_IDS = ["{}-{}".format(case.name, mode) for case, mode in CASE_MODE_PAIRS]


def oracle_output(stencil, *args, **kwargs):
    from oracle.cref import reference_c
    return reference_c(stencil, *args, **kwargs)


def periodic_oracle(make, grid, *tables):
    """Oracle for boundary_handling='periodic' (an addition; the reference's 'wrap' is 'zero'):
    pad the grid by the ghost depth with numpy.pad(mode='wrap'), run the restated reference C in
    'zero' mode (interior only) on the padded grid, crop.  ``make(mode)`` builds the stencil."""
    stencil = make('zero')
    ghost = stencil.ghost_depth
    padded = numpy.ascontiguousarray(numpy.pad(grid, [(g, g) for g in ghost], mode='wrap'))
    result = oracle_output(stencil, padded, *tables)
    return numpy.ascontiguousarray(result[tuple(slice(g, g + n) for g, n in zip(ghost, grid.shape))])
