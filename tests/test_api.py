"""API behaviour mirrored from the reference's unit tests (test/test_stencil_kernel.py,
test/test_stencil_backend.py, test/test_neighborhood.py, test/test_halo_enumerator.py,
test/test_boundary_handling.py:13-46) -- no device needed."""
import itertools

import numpy
import pytest

from stencil_code_b200 import Stencil, StencilException, Neighborhood, HaloEnumerator, product
from stencil_code_b200.backend.stencil_backend import StencilBackend
from stencil_code_b200.library.jacobi_stencil import Jacobi


def test_reference_import_paths_still_work():
    from stencil_code.stencil_kernel import Stencil as S
    from stencil_code.neighborhood import Neighborhood as N
    from stencil_code.library.laplacian import LaplacianKernel
    assert S is Stencil and N is Neighborhood and issubclass(LaplacianKernel, Stencil)


def test_no_kernel_exception():
    class BadKernel(Stencil):
        pass
    with pytest.raises(StencilException):
        BadKernel()


def test_bad_boundary_handling():
    class Kernel(Stencil):
        neighborhoods = [[(0, 0)]]

        def kernel(self, *args):
            pass
    with pytest.raises(StencilException) as context:
        Kernel(boundary_handling='dog')
    assert "boundary handling value" in context.value.args[0]


def test_no_neighborhood():
    class NoNeighborhood(Stencil):
        def kernel(self):
            pass
    with pytest.raises(StencilException) as context:
        NoNeighborhood()
    assert "neighborhoods must be defined" in context.value.args[0]


def test_bad_neighborhood_id():
    class OneNeighborhood(Stencil):
        neighborhoods = [[(0, 0)]]

        def kernel(self):
            pass
    with pytest.raises(StencilException) as context:
        next(OneNeighborhood().neighbors((0, 1), 2))
    assert "Undefined neighborhood identifier" in context.value.args[0]


def test_default_distance():
    jacobi = Jacobi()
    assert jacobi.distance((0, 0), (10, 0)) == 10.0
    assert jacobi.distance((3, 0), (0, 4)) == 5.0


def test_backends_that_need_ctree_raise(monkeypatch):
    """strict mode (the default of ``stencil_code_b200``; importing the ``stencil_code`` alias
    package maps the reference's backend names to 'cuda' instead, tests/test_reference_crosscheck.py)"""
    monkeypatch.setenv("STENCIL_CUDA_LEGACY_BACKENDS", "raise")
    for name in ('c', 'omp', 'ocl', 'opencl', 'nope'):
        with pytest.raises(StencilException):
            Jacobi(backend=name)
    monkeypatch.setenv("STENCIL_CUDA_LEGACY_BACKENDS", "cuda")
    with pytest.raises(StencilException):
        Jacobi(backend='nope')
    assert Jacobi(backend='omp').backend_name == 'cuda'


def test_backend_base_requires_visit_interior_points_loop():
    class NoInteriorPoints(StencilBackend):
        pass
    with pytest.raises(StencilException) as context:
        NoInteriorPoints(kernel=Jacobi())
    assert "must define a visit_InteriorPointsLoop method" in context.value.args[0]


def test_python_clamping_neighbors():               # test_boundary_handling.py:13-46
    class Clamper(Stencil):
        neighborhoods = [Neighborhood.von_neuman_neighborhood(radius=1, dim=2)]

        def kernel(self, in_grid, out_grid):
            pass
    clamper = Clamper(backend='python', boundary_handling='clamp')
    clamper.current_shape = (10, 10)
    assert list(clamper.neighbors((9, 9), 0)) == [(8, 9), (9, 8), (9, 9), (9, 9), (9, 9)]
    assert list(clamper.neighbors((0, 0), 0)) == [(0, 0), (0, 0), (0, 0), (0, 1), (1, 0)]
    assert list(clamper.neighbors((3, 3), 0)) == [(2, 3), (3, 2), (3, 3), (3, 4), (4, 3)]


def test_neighborhoods():
    assert Neighborhood.origin_of_dim(3) == (0, 0, 0)
    assert sorted(Neighborhood.von_neuman_neighborhood()) == sorted([(-1, 0), (0, -1), (0, 1), (1, 0), (0, 0)])
    assert len(Neighborhood.von_neuman_neighborhood(3, 2)) == 25
    assert Neighborhood.moore_neighborhood(1, 2)[0] == (-1, -1)           # order is part of the contract
    assert Neighborhood.von_neuman_neighborhood(1, 3, include_origin=False) == [
        (-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0), (1, 0, 0)]
    n, c, h = Neighborhood.compute_from_indices([[0, 1, 0], [1, 4, 8], [0, 8, 0]])
    assert len(n) == 5 and [int(v) for v in c] == [1, 1, 4, 8, 8] and h[0][0] == 1 and h[1][1] == 1
    assert sorted(n) == sorted(Neighborhood.von_neuman_neighborhood(1, 2))


def test_halo_enumerator():
    with pytest.raises(AssertionError) as context:
        HaloEnumerator([1, 1], [5, 5, 5])
    assert "HaloEnumerator halo" in context.value.args[0]
    assert sorted(HaloEnumerator([2], [5])) == [(0,), (1,), (3,), (4,)]
    for dimension in range(1, 5):
        for halo_size in range(1, 3):
            shape = [5] * dimension

            class Kernel(Stencil):
                neighborhoods = [Neighborhood.moore_neighborhood(radius=halo_size, dim=dimension,
                                                                 include_origin=False)]

                def kernel(self):
                    pass
            kernel = Kernel(backend='python', boundary_handling='zero')
            everything = set(itertools.product(*[range(5)] * dimension))
            interior = set(kernel.interior_points(numpy.zeros(shape)))
            halo = list(kernel.halo_points(numpy.zeros(shape)))
            assert len(halo) == len(set(halo)) == product(shape) - len(interior)
            assert not interior & set(halo) and interior | set(halo) == everything


def test_iterations_keyword_and_iterate_on_the_python_backend():
    """additive API: stencil(grid, iterations=n) == stencil.iterate(grid, n) == n chained calls"""
    import numpy
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    grid = numpy.random.RandomState(9).random_sample((5, 6, 7)).astype(numpy.float32)
    stencil = Jacobi3D(0.5, 0.0625, backend='python')
    chained = stencil(stencil(stencil(grid)))
    assert (stencil(grid, iterations=3) == chained).all() and (stencil.iterate(grid, 3) == chained).all()
    assert (stencil.iterate(grid, 0) == grid).all()
    flat = numpy.random.RandomState(9).random_sample((9, 11)).astype(numpy.float32)
    conv = ConvolutionFilter(numpy.array([[0, 0.25, 0], [0.25, 0, 0.25], [0, 0.25, 0]]), backend='python')
    assert (conv(flat, iterations=2) == conv(conv(flat))).all() and (conv.iterate(flat, 2) == conv(conv(flat))).all()


def test_jacobi_example_script_runs_on_the_python_backend():
    """examples/jacobi_3d.py (the reference's stale example restated on the current API)"""
    import numpy
    from examples import jacobi_3d
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    result = jacobi_3d.main(["--backend", "python", "--shape", "5", "6", "7", "--iterations", "2"])
    grid = numpy.random.RandomState(1).random_sample((5, 6, 7)).astype(numpy.float32)
    stencil = Jacobi3D(alpha=0.4, beta=0.1, backend='python')
    assert (result == stencil(stencil(grid))).all()
