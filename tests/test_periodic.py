"""boundary_handling='periodic': true wrap-around, an explicit addition (SURVEY section 8, table B:
the reference's 'wrap' is a typo-disabled flag and behaves like 'zero', which 'wrap' keeps doing)."""
import numpy
import pytest

from cases import COEFF27, CONV5, OddWeights, Stencil1d, WideStar3D, periodic_oracle
from stencil_code_b200 import StencilException
from stencil_code_b200.library.convolution import ConvolutionFilter
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.laplacian import LaplacianKernel
from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27

CASES = {
    "laplacian": (lambda backend, mode: LaplacianKernel(backend=backend, boundary_handling=mode), (9, 12, 40), ()),
    "jacobi3d": (lambda backend, mode: Jacobi3D(0.5, 0.0625, backend=backend, boundary_handling=mode), (16, 20, 132), ()),
    "widestar": (lambda backend, mode: WideStar3D(backend=backend, boundary_handling=mode), (7, 9, 33), ()),
    "laplacian27": (lambda backend, mode: SpecializedLaplacian27(backend=backend, boundary_handling=mode),
                    (8, 11, 36), (COEFF27,)),
    "odd2d": (lambda backend, mode: OddWeights(backend=backend, boundary_handling=mode), (37, 260), ()),
    "conv5": (lambda backend, mode: ConvolutionFilter(CONV5, backend=backend, boundary_handling=mode), (21, 64), ()),
    "1d": (lambda backend, mode: Stencil1d(backend=backend, boundary_handling=mode), (50,), ()),
}


def bits(a):
    a = numpy.ascontiguousarray(a)
    return a.view(numpy.uint32 if a.dtype == numpy.float32 else numpy.uint64)


def test_wrap_still_means_zero_and_periodic_is_separate():
    grid = numpy.random.RandomState(4).random_sample((6, 7)).astype(numpy.float32)
    wrap = OddWeights(backend='python', boundary_handling='wrap')(grid)
    zero = OddWeights(backend='python', boundary_handling='zero')(grid)
    periodic = OddWeights(backend='python', boundary_handling='periodic')(grid)
    assert (bits(wrap) == bits(zero)).all() and (wrap[0] == 0).all()
    assert (periodic[0] != 0).all()
    with pytest.raises(StencilException):
        OddWeights(backend='python', boundary_handling='torus')


@pytest.mark.parametrize("name", ["laplacian", "odd2d", "1d", "widestar"])
def test_python_backend_wraps_neighbours(name):
    make, shape, tables = CASES[name]
    shape = tuple(min(n, 12) for n in shape)
    grid = numpy.random.RandomState(5).random_sample(shape).astype(numpy.float32)
    actual = make('python', 'periodic')(grid, *tables)
    ghost = make('python', 'zero').ghost_depth
    padded = numpy.pad(grid, [(g, g) for g in ghost], mode='wrap')
    expected = make('python', 'zero')(padded, *tables)[tuple(slice(g, g + n) for g, n in zip(ghost, shape))]
    assert (bits(actual) == bits(expected)).all()


def test_periodic_takes_the_point_kernel_and_wraps_indices(monkeypatch):
    from test_codegen_cpu import make_plan
    from stencil_code_b200.backend.cuda import runtime
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    plan = make_plan(LaplacianKernel(boundary_handling='periodic'), [((64, 128, 256), numpy.float32)])
    assert plan.strategy == 'direct' and "SC_WRAP(iz + (1), 64)" in plan.source and "SC_CLAMP(i" not in plan.source
    monkeypatch.setenv("STENCIL_CUDA_DIRECT", "flat")          # the one-point-per-thread form of the same kernel
    flat = make_plan(LaplacianKernel(boundary_handling='periodic'), [((64, 128, 256), numpy.float32)])
    assert "SC_WRAP(i0 + (1), 64)" in flat.source
    runtime.compile_source(flat.source, "periodic_flat.cu", flat.options)
    monkeypatch.delenv("STENCIL_CUDA_DIRECT")
    runtime.compile_source(plan.source, "periodic.cu", plan.options)
    with pytest.raises(StencilException):
        make_plan(LaplacianKernel(boundary_handling='periodic'), [((64, 128, 256), numpy.float32)], 'stream')
    with pytest.raises(StencilException):                      # reach 2 in a dimension of extent 1
        make_plan(WideStar3D(boundary_handling='periodic'), [((4, 1, 64), numpy.float32)])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_periodic_matches_the_padded_oracle(name, monkeypatch):
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    make, shape, tables = CASES[name]
    grid = (numpy.random.RandomState(6).random_sample(shape) * 4).astype(numpy.float32)
    stencil = make('cuda', 'periodic')
    once = periodic_oracle(lambda mode: make('python', mode), grid, *tables)
    assert (bits(stencil(grid, *tables)) == bits(once)).all()
    if name != "conv5":
        thrice = once
        for _ in range(2):
            thrice = periodic_oracle(lambda mode: make('python', mode), thrice, *tables)
        assert (bits(stencil.iterate(grid, 3, *tables)) == bits(thrice)).all()


@pytest.mark.gpu
def test_cuda_periodic_large_grid_through_host_and_device_paths():
    import torch
    make = lambda backend, mode: Jacobi3D(0.5, 0.0625, backend=backend, boundary_handling=mode)
    grid = numpy.random.RandomState(7).random_sample((96, 256, 512)).astype(numpy.float32)     # 48 MiB: pipelining size
    expected = periodic_oracle(lambda mode: make('python', mode), grid)
    stencil = make('cuda', 'periodic')
    assert (bits(stencil(grid)) == bits(expected)).all()
    assert (bits(stencil(torch.from_numpy(grid).cuda()).cpu().numpy()) == bits(expected)).all()
