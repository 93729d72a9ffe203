"""fuse(a, b, ...): composition of stencils with the grid resident on the device, pairs of
star-in-z stencils in one launch (reference: ipython_notebooks/fusing_stencils, `composable`)."""
import numpy
import pytest

from cases import WideStar3D, oracle_output
from stencil_code_b200 import fuse, StencilException
from stencil_code_b200.library.jacobi_3d import Jacobi3D
from stencil_code_b200.library.laplacian import LaplacianKernel
from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27


def bits(a):
    return numpy.ascontiguousarray(a).view(numpy.uint32)


def test_python_backend_composition_is_plain_chaining():
    grid = numpy.random.RandomState(3).random_sample((6, 7, 8)).astype(numpy.float32)
    a, b = LaplacianKernel(backend='python'), Jacobi3D(1.0, 0.25, backend='python')
    assert (bits(fuse(a, b)(grid)) == bits(b(a(grid)))).all()
    assert (bits(fuse(a)(grid)) == bits(a(grid))).all()


def test_fuse_rejects_mixed_or_unknown_backends():
    with pytest.raises(StencilException):
        fuse(LaplacianKernel(backend='python'), Jacobi3D(backend='cuda'))
    with pytest.raises(StencilException):
        fuse()


def test_pair_eligibility_and_compile():
    from test_codegen_cpu import make_plan
    from stencil_code_b200.backend.cuda import codegen_temporal, planner, runtime
    shape = (32, 64, 512)
    for mode in ('clamp', 'zero', 'copy'):
        first = make_plan(WideStar3D(boundary_handling=mode), [(shape, numpy.float32)], 'stream')
        second = make_plan(Jacobi3D(boundary_handling=mode), [(shape, numpy.float32)], 'stream')
        for a, b in ((first, second), (second, first)):
            ok, why = codegen_temporal.pair_eligible(a.program, a.domain, b.program, b.domain)
            assert ok, why
            plan = codegen_temporal.make_plan(a.program, a.domain, planner.B200, a.options,
                                              second=b.program, second_domain=b.domain)
            assert plan.notes['config']['HY'] == 2 and plan.notes['config']['OUT_ROWS'] == plan.notes['config']['MY'] - 4
            assert ("halo_mask2" in plan.source) == (mode != 'clamp')      # two ghost depths, two halos
            runtime.compile_source(plan.source, "pair.cu", plan.options)
    other_mode = make_plan(Jacobi3D(boundary_handling='zero'), [(shape, numpy.float32)], 'stream')
    assert not codegen_temporal.pair_eligible(first.program, first.domain, other_mode.program, other_mode.domain)[0]
    with_table = make_plan(SpecializedLaplacian27(boundary_handling='copy'),
                           [(shape, numpy.float32), ((4,), numpy.float32)], 'stream')
    assert not codegen_temporal.pair_eligible(first.program, first.domain, with_table.program, with_table.domain)[0]


# ---- on the GPU ------------------------------------------------------------------------------------
def chain(stencils, grid, tables=None):
    tables = tables or [()] * len(stencils)
    for stencil, extra in zip(stencils, tables):
        grid = oracle_output(stencil, grid, *extra)
    return grid


FUSE_SHAPES = [(12, 70, 256), (19, 41, 388), (9, 31, 516)]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["clamp", "zero", "copy"])
@pytest.mark.parametrize("shape", FUSE_SHAPES, ids=["x".join(map(str, s)) for s in FUSE_SHAPES])
def test_fused_pairs_match_the_oracle_chain(mode, shape, monkeypatch):
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")       # small planes: not "worthwhile"
    grid = numpy.random.RandomState(29).random_sample(shape).astype(numpy.float32)
    makers = [lambda backend: LaplacianKernel(backend=backend, boundary_handling=mode),
              lambda backend: Jacobi3D(0.5, 0.125, backend=backend, boundary_handling=mode),
              lambda backend: WideStar3D(backend=backend, boundary_handling=mode)]
    for order in ((0, 1), (1, 0), (2, 1), (1, 2), (2, 2), (0, 1, 2), (2, 0, 1, 1)):
        cuda = fuse(*[makers[k]('cuda') for k in order])
        expected = chain([makers[k]('python') for k in order], grid)
        actual = cuda(grid)
        assert cuda.last_schedule == ['pair'] * (len(order) // 2) + ['single'] * (len(order) % 2)
        different = bits(actual) != bits(expected)
        assert not different.any(), "{} {}: {} differ, first at {}".format(
            order, mode, int(different.sum()), tuple(numpy.argwhere(different)[0]))


@pytest.mark.gpu
def test_fusion_falls_back_to_single_launches(monkeypatch):
    """Stages with tables, rank-2 stencils, STENCIL_CUDA_TEMPORAL=0: one launch per stage, the grid
    still stays on the device; torch tensors in, torch tensor out, the input untouched."""
    import torch
    from cases import COEFF27, OddWeights
    grid = numpy.random.RandomState(31).random_sample((10, 40, 256)).astype(numpy.float32)
    stencils = [SpecializedLaplacian27(backend='cuda', boundary_handling='copy'), Jacobi3D(0.5, 0.125, backend='cuda')]
    pythons = [SpecializedLaplacian27(backend='python', boundary_handling='copy'), Jacobi3D(0.5, 0.125, backend='python')]
    fused = fuse(*stencils)
    actual = fused(grid, tables=[(COEFF27,), ()])
    assert fused.last_schedule == ['single', 'single']
    assert (bits(actual) == bits(chain(pythons, grid, [(COEFF27,), ()]))).all()
    tensor = torch.from_numpy(grid).cuda()
    result = fused(tensor, tables=[(torch.from_numpy(COEFF27).cuda(),), ()])
    assert result.is_cuda and (bits(result.cpu().numpy()) == bits(actual)).all()
    assert (bits(tensor.cpu().numpy()) == bits(grid)).all()

    flat = numpy.random.RandomState(32).random_sample((300, 516)).astype(numpy.float32)
    pair2d = fuse(OddWeights(backend='cuda'), OddWeights(backend='cuda'), OddWeights(backend='cuda'))
    expected = chain([OddWeights(backend='python')] * 3, flat)
    assert (bits(pair2d(flat)) == bits(expected)).all() and pair2d.last_schedule == ['single'] * 3

    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "0")
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    plain = fuse(LaplacianKernel(backend='cuda'), Jacobi3D(0.5, 0.125, backend='cuda'))
    expected = chain([LaplacianKernel(backend='python'), Jacobi3D(0.5, 0.125, backend='python')], grid)
    assert (bits(plain(grid)) == bits(expected)).all() and plain.last_schedule == ['single', 'single']


@pytest.mark.gpu
def test_fused_pair_on_a_large_grid_equals_two_calls(monkeypatch):
    """512 x 1024 x 1024 (worthwhile tiles): one fused launch == two single launches, compared on
    the device; and it is the faster of the two."""
    import torch
    monkeypatch.delenv("STENCIL_CUDA_TEMPORAL", raising=False)
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    first, second = LaplacianKernel(backend='cuda'), Jacobi3D(1.0, 0.25, backend='cuda')
    grid = torch.rand((256, 1024, 1024), device='cuda')
    fused = fuse(first, second)
    together = fused(grid)
    assert fused.last_schedule == ['pair']
    apart = second(first(grid))
    assert torch.equal(together.view(torch.int32), apart.view(torch.int32))

    def timed(function):
        function(); torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(5):
            function()
        stop.record(); torch.cuda.synchronize()
        return start.elapsed_time(stop) / 5
    assert timed(lambda: fused(grid)) < timed(lambda: second(first(grid)))
