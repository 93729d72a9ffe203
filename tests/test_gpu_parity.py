"""GPU parity: backend='cuda' against the CPU oracle (restated reference C), bit for bit.

Every comparison is on the FULL grid (halo included) and on bit patterns: float32 results are
compared as uint32.  Float tolerance used anywhere in this file: none -- summation order is
preserved and FMA contraction is off, so the north-star's "bit-exact" branch applies to every
case here (the <= 4 ulp branch is not needed).
"""
import os

import numpy
import pytest

from cases import CASES, CASE_MODE_PAIRS, PAIR_IDS, oracle_output
from stencil_code_b200.stencil_exception import StencilException

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STRATEGIES = ("direct", "auto")        # for the small reference-sized grids


def bits(array):
    array = numpy.ascontiguousarray(array)
    return array.view(numpy.uint32 if array.dtype == numpy.float32 else numpy.uint64)


def assert_bit_identical(actual, expected):
    assert actual.shape == expected.shape and actual.dtype == expected.dtype
    different = bits(actual) != bits(expected)
    if different.any():
        where = numpy.argwhere(different)
        first = tuple(where[0])
        raise AssertionError("{} of {} elements differ; first at {}: got {!r}, expected {!r}".format(
            int(different.sum()), different.size, first, actual[first], expected[first]))


def set_strategy(monkeypatch, strategy):
    if strategy == "auto":
        monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    else:
        monkeypatch.setenv("STENCIL_CUDA_STRATEGY", strategy)


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("case,mode", CASE_MODE_PAIRS, ids=PAIR_IDS)
def test_small_grids_match_oracle(case, mode, strategy, monkeypatch):
    set_strategy(monkeypatch, strategy)
    grid = case.grid(seed=1)
    expected = oracle_output(case.make(mode, 'python'), grid, *case.extra)
    actual = case.make(mode, 'cuda')(grid, *case.extra)
    assert isinstance(actual, numpy.ndarray)
    assert_bit_identical(actual, expected)


@pytest.mark.parametrize("case,mode", [(c, m) for c, m in CASE_MODE_PAIRS if c.golden],
                         ids=["{}-{}".format(c.name, m) for c, m in CASE_MODE_PAIRS if c.golden])
def test_reference_made_fixtures(case, mode):
    """Integer-valued fixtures computed by the unmodified reference: exact equality (interior only
    for distance-weighted clamp cases, see tests/test_oracle_golden.py)."""
    from test_oracle_golden import comparable, INEXACT
    data = numpy.load(os.path.join(GOLDEN, "{}__{}.npz".format(case.name, mode)))
    stencil = case.make(mode, 'cuda')
    actual = comparable(case, mode, stencil, stencil(data["exact_in"], *case.extra))
    expected = comparable(case, mode, stencil, data["exact_out"])
    if case.name in INEXACT:
        scale = float(numpy.abs(expected).max())
        numpy.testing.assert_array_almost_equal(actual / scale, expected / scale, decimal=6)
    else:
        numpy.testing.assert_array_equal(actual, expected)


# shapes that exercise tile edges, unaligned rows, thin grids, and TMA-eligible rows
SHAPE_SWEEP = {
    "laplacian": [(5, 5, 5), (3, 4, 7), (33, 20, 132), (16, 40, 256), (70, 37, 64), (4, 130, 136)],
    "jacobi3d": [(40, 24, 128), (19, 65, 260)],
    "jacobi2d": [(128, 128), (130, 260), (1024, 1024)],       # reference test/test_c_end_to_end.py:72-74 (128^2)
    "simple_moore": [(5, 5), (102, 7), (64, 128), (130, 260), (37, 1024), (300, 516)],
    "convolution5_modes": [(8, 8), (64, 128), (131, 388), (50, 1028)],
    "diagnostic": [(8, 8), (96, 256), (45, 132)],
    "two_d_heat": [(16, 16, 16), (9, 34, 128)],
    "laplacian27": [(32, 32, 32), (10, 33, 136)],
    "cross3d_r2": [(12, 20, 128), (9, 9, 12)],
    "odd_weights": [(64, 128), (33, 36)],
    "stencil1d": [(8,), (1000,), (4096,)],
    "better_bilateral": [(64, 32), (40, 128)],
    "bilateral_r9": [(70, 260), (33, 1028)],
}
SWEEP = [(case, mode, shape) for case in CASES if case.name in SHAPE_SWEEP
         for mode in case.modes for shape in SHAPE_SWEEP[case.name]]


@pytest.mark.parametrize("strategy", ("stream", "auto"))
@pytest.mark.parametrize("case,mode,shape", SWEEP,
                         ids=["{}-{}-{}".format(c.name, m, "x".join(map(str, s))) for c, m, s in SWEEP])
def test_shape_sweep_matches_oracle(case, mode, shape, strategy, monkeypatch):
    """'stream' forces the TMA streaming / tiled generators even on grids the planner would hand
    to 'direct' (small, L2 resident); shapes it cannot take (odd rows, rank 1) are skipped."""
    set_strategy(monkeypatch, strategy)
    grid = case.grid(seed=2, shape=shape)
    expected = oracle_output(case.make(mode, 'python'), grid, *case.extra)
    try:
        actual = case.make(mode, 'cuda')(grid, *case.extra)
    except StencilException as error:
        if strategy == "stream" and "cannot handle" in str(error):
            pytest.skip(str(error))
        raise
    assert_bit_identical(actual, expected)


def test_float64_grid():                       # reference test_boundary_handling.py:56 uses float64
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
    grid = numpy.random.RandomState(5).random_sample((20, 24)) * 1000
    for mode in ('clamp', 'zero', 'copy'):
        expected = oracle_output(DiagnosticStencil(backend='python', boundary_handling=mode), grid)
        actual = DiagnosticStencil(backend='cuda', boundary_handling=mode)(grid)
        assert actual.dtype == numpy.float64
        assert_bit_identical(actual, expected)


def test_strict_double_mode_equals_demoted_mode(monkeypatch):
    """The exact float32 demotion must not change a bit (normal-range inputs)."""
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.two_d_heat import TwoDHeatFlow
    for make, shape in ((lambda: Jacobi3D(1.0, 0.25), (12, 16, 128)),
                        (lambda: TwoDHeatFlow(), (8, 16, 128))):
        grid = (numpy.random.RandomState(7).random_sample(shape) * 1000).astype(numpy.float32)
        monkeypatch.delenv("STENCIL_CUDA_STRICT", raising=False)
        demoted = make()(grid)
        monkeypatch.setenv("STENCIL_CUDA_STRICT", "1")
        strict = make()(grid)
        assert_bit_identical(demoted, strict)
        assert_bit_identical(demoted, oracle_output(make(), grid))


def test_torch_tensor_in_torch_tensor_out():
    import torch
    from stencil_code_b200.library.laplacian import LaplacianKernel
    grid = (numpy.random.RandomState(9).random_sample((16, 24, 128)) * 1000).astype(numpy.float32)
    expected = oracle_output(LaplacianKernel(backend='python'), grid)
    result = LaplacianKernel()(torch.from_numpy(grid).cuda())
    assert isinstance(result, torch.Tensor) and result.is_cuda
    assert_bit_identical(result.cpu().numpy(), expected)


def test_iterate_equals_repeated_calls():
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    grid = numpy.random.RandomState(11).random_sample((24, 16, 128)).astype(numpy.float32)
    stencil = Jacobi3D(1.0, 0.25)
    looped = grid
    for _ in range(5):
        looped = stencil(looped)
    assert_bit_identical(stencil.iterate(grid, 5), looped)
    assert_bit_identical(stencil(grid, iterations=5), looped)         # the keyword form of the same
    concrete = stencil.specializer.specialize((grid,))
    assert_bit_identical(concrete(grid, iterations=5), looped)
    reference = grid
    python_side = Jacobi3D(1.0, 0.25, backend='python')
    for _ in range(5):
        reference = oracle_output(python_side, reference)
    assert_bit_identical(looped, reference)


def test_wrong_shape_raises_type_error():
    from stencil_code_b200.library.laplacian import LaplacianKernel
    stencil = LaplacianKernel()
    stencil(numpy.zeros((8, 8, 8), numpy.float32))
    concrete = stencil.specializer.specialize((numpy.zeros((8, 8, 8), numpy.float32),))
    with pytest.raises(TypeError):
        concrete(numpy.zeros((8, 8, 9), numpy.float32))


def test_host_side_torch_tensors_are_refused_everywhere():
    import torch
    from stencil_code_b200 import StencilException, fuse
    from stencil_code_b200.library.laplacian import LaplacianKernel
    stencil = LaplacianKernel()
    host = torch.zeros((8, 8, 8))
    for call in (lambda: stencil(host), lambda: stencil.iterate(host, 2), lambda: fuse(stencil, stencil)(host)):
        with pytest.raises(StencilException, match="CUDA device"):
            call()


def test_product_never_imports_the_oracle():
    import sys
    import subprocess
    code = ("import sys, numpy; sys.path.insert(0, '.');"
            "from stencil_code_b200.library.laplacian import LaplacianKernel;"
            "LaplacianKernel()(numpy.ones((8, 8, 8), numpy.float32));"
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], cwd=root, check=True)


def test_pipelined_host_path_matches_oracle():
    """Grids above the pipelining threshold are swept in axis-0 chunks on three streams."""
    import stencil_code_b200
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from stencil_code_b200.library.simple import SimpleKernel
    from stencil_code_b200.backend.cuda.runtime import ConcreteCudaStencil
    for make, shape in ((lambda backend, mode: LaplacianKernel(backend=backend, boundary_handling=mode), (96, 256, 512)),
                        (lambda backend, mode: SimpleKernel(backend=backend, boundary_handling=mode), (4096, 4096))):
        assert numpy.prod(shape) * 4 >= ConcreteCudaStencil.PIPELINE_MIN_BYTES
        pinned = stencil_code_b200.pinned_empty(shape)
        pinned[...] = (numpy.random.RandomState(13).random_sample(shape) * 1000).astype(numpy.float32)
        for mode in ('clamp', 'copy'):
            expected = oracle_output(make('python', mode), numpy.array(pinned), variant='c_parallel')
            assert_bit_identical(numpy.array(make('cuda', mode)(pinned)), expected)
            assert_bit_identical(numpy.array(make('cuda', mode)(numpy.array(pinned))), expected)   # pageable


TEMPORAL_SHAPES = [(12, 70, 256), (19, 64, 388), (33, 100, 260), (9, 31, 512)]


@pytest.mark.parametrize("mode", ["clamp", "zero", "copy"])
@pytest.mark.parametrize("shape", TEMPORAL_SHAPES, ids=["x".join(map(str, s)) for s in TEMPORAL_SHAPES])
def test_temporal_blocking_equals_separate_sweeps(mode, shape, monkeypatch):
    """iterate() fuses pairs of sweeps into one kernel (codegen_temporal); the result must be
    bit-identical to sweeping n times -- checked against the CPU oracle applied n times."""
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.laplacian import LaplacianKernel
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")      # small planes: not "worthwhile"
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    for make in (lambda backend: Jacobi3D(1.0, 0.25, backend=backend, boundary_handling=mode),
                 lambda backend: LaplacianKernel(backend=backend, boundary_handling=mode)):
        grid = numpy.random.RandomState(17).random_sample(shape).astype(numpy.float32)
        stencil = make('cuda')
        concrete = stencil.specializer.specialize((grid,))
        assert concrete.temporal() is not None, "temporal kernel expected for this stencil"
        python_side = make('python')
        expected = grid
        for sweeps in (1, 2, 3, 4, 5):
            expected = oracle_output(python_side, expected)
            if sweeps in (2, 3, 5):
                assert_bit_identical(stencil.iterate(grid, sweeps), expected)
    import torch
    tensor = torch.from_numpy(grid).cuda()
    assert_bit_identical(stencil.iterate(tensor, 4).cpu().numpy(),
                         oracle_output(python_side, oracle_output(python_side, oracle_output(
                             python_side, oracle_output(python_side, grid)))))
    assert_bit_identical(tensor.cpu().numpy(), grid)          # the caller's tensor is not modified


LONG_RUNS = [("direct", (40, 48, 64), 60), ("stream", (128, 128, 128), 40), ("stream", (24, 260, 512), 31),
             ("direct", (300, 260), 60), ("stream", (1024, 1024), 40), ("stream", (700, 2052), 33)]


@pytest.mark.parametrize("strategy,shape,sweeps", LONG_RUNS,
                         ids=["{}-{}-x{}".format(s, "x".join(map(str, shape)), n) for s, shape, n in LONG_RUNS])
def test_long_iterated_runs_with_programmatic_dependent_launch(strategy, shape, sweeps, monkeypatch):
    """iterate(n) enqueues its sweeps from C with programmatic stream serialization: sweep k+1's CTAs
    become resident while sweep k drains and hold at griddepcontrol.wait.  Small grids, where
    several sweeps are in flight at once, are where an ordering bug would show: n sweeps must
    equal the oracle applied n times, and the same run with the attribute off."""
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from cases import OddWeights
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", strategy)
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "0")
    make = (lambda backend: Jacobi3D(0.5, 0.0625, backend=backend)) if len(shape) == 3 else \
        (lambda backend: OddWeights(backend=backend, boundary_handling='copy'))
    grid = numpy.random.RandomState(23).random_sample(shape).astype(numpy.float32)
    python_side = make('python')
    expected = grid
    for _ in range(sweeps):
        expected = oracle_output(python_side, expected)
    stencil = make('cuda')
    assert stencil.specializer.specialize((grid,)).plan.strategy == strategy
    for _ in range(3):                                  # timing-dependent bugs get three chances
        assert_bit_identical(stencil.iterate(grid, sweeps), expected)
    monkeypatch.setenv("STENCIL_CUDA_PDL", "0")
    assert_bit_identical(stencil.iterate(grid, sweeps), expected)


ODD_WIDTHS = [((20, 40, 513), "jacobi3d"), ((9, 70, 258), "jacobi3d"), ((12, 33, 387), "laplacian27"),
              ((300, 1025), "odd2d"), ((130, 515), "conv5")]


@pytest.mark.parametrize("mode", ["clamp", "zero", "copy"])
@pytest.mark.parametrize("shape,name", ODD_WIDTHS, ids=["{}-{}".format(n, "x".join(map(str, s))) for s, n in ODD_WIDTHS])
def test_iterate_on_rows_that_are_no_multiple_of_16_bytes(shape, name, mode, monkeypatch):
    """TMA cannot address such grids in place; iterate() packs them into padded rows once, sweeps
    there (streaming and fused kernels) and unpacks: same bits as the oracle applied n times."""
    import torch
    from cases import COEFF27, CONV5, OddWeights
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")
    make = {"jacobi3d": lambda backend: Jacobi3D(0.5, 0.0625, backend=backend, boundary_handling=mode),
            "laplacian27": lambda backend: SpecializedLaplacian27(backend=backend, boundary_handling=mode),
            "odd2d": lambda backend: OddWeights(backend=backend, boundary_handling=mode),
            "conv5": lambda backend: ConvolutionFilter(CONV5 / 32.0, backend=backend, boundary_handling=mode)}[name]
    tables = (COEFF27,) if name == "laplacian27" else ()
    grid = numpy.random.RandomState(37).random_sample(shape).astype(numpy.float32)
    stencil, python_side = make('cuda'), make('python')
    hidden = (stencil.coefficients,) if name == "conv5" else ()   # ConvolutionFilter passes it itself
    concrete = stencil.specializer.specialize((grid,) + tables + hidden)
    assert concrete.plan.strategy == 'direct'                     # single calls: dense rows, no TMA
    twin = concrete._pitched_twin()
    assert twin is not None and twin.plan.strategy == 'stream' and twin.plan.domain.row_pitch % 4 == 0
    expected = grid
    for sweeps in range(1, 6):
        expected = oracle_output(python_side, expected, *tables)
        if sweeps in (1, 2, 5):
            assert_bit_identical(stencil.iterate(grid, sweeps, *tables), expected)
    tensor = torch.from_numpy(grid).cuda()
    result = stencil.iterate(tensor, 5, *[torch.from_numpy(t).cuda() for t in tables])
    assert result.shape == tensor.shape and result.is_contiguous()
    assert_bit_identical(result.cpu().numpy(), expected)
    assert_bit_identical(tensor.cpu().numpy(), grid)
    assert_bit_identical(stencil(grid, *tables), oracle_output(python_side, grid, *tables))   # single call unchanged


def test_padded_rows_are_not_used_where_they_do_not_apply(monkeypatch):
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    stencil = Jacobi3D(0.5, 0.0625, backend='cuda')
    aligned = numpy.zeros((8, 16, 64), numpy.float32)
    assert stencil.specializer.specialize((aligned,))._pitched_twin() is None
    double = numpy.random.RandomState(2).random_sample((8, 16, 67))
    concrete = stencil.specializer.specialize((double,))
    assert concrete._pitched_twin() is None
    python_side = Jacobi3D(0.5, 0.0625, backend='python')
    expected = oracle_output(python_side, oracle_output(python_side, double))
    assert (stencil.iterate(double, 2).view(numpy.uint64) == expected.view(numpy.uint64)).all()
    monkeypatch.setenv("STENCIL_CUDA_PITCHED", "0")
    odd = numpy.random.RandomState(3).random_sample((8, 16, 67)).astype(numpy.float32)
    assert Jacobi3D(0.5, 0.0625, backend='cuda').specializer.specialize((odd,))._pitched_twin() is None


@pytest.mark.parametrize("mode", ["clamp", "zero", "copy", "periodic"])
def test_point_kernel_flat_and_tiled_forms_agree(mode, monkeypatch):
    """'direct' walks rank-2/3 grids with RZ x RY points per thread (shared loads) and everything
    else with one point per thread; both forms against the oracle on the same grids, float32 and
    float64, ragged tile edges (rows and planes that are no multiples of RY / RZ)."""
    from cases import OddWeights, WideStar3D, periodic_oracle
    from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
    from cases import COEFF27
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "direct")
    cases = [(lambda backend, m: WideStar3D(backend=backend, boundary_handling=m), (11, 30, 132), (), numpy.float32),
             (lambda backend, m: SpecializedLaplacian27(backend=backend, boundary_handling=m), (7, 13, 100), (COEFF27,), numpy.float32),
             (lambda backend, m: OddWeights(backend=backend, boundary_handling=m), (51, 261), (), numpy.float32),
             (lambda backend, m: WideStar3D(backend=backend, boundary_handling=m), (9, 10, 97), (), numpy.float64)]
    for make, shape, tables, dtype in cases:
        grid = numpy.random.RandomState(41).random_sample(shape).astype(dtype)
        if mode == "periodic":
            expected = periodic_oracle(lambda m: make('python', m), grid, *tables)
        else:
            expected = oracle_output(make('python', mode), grid, *tables)
        view = numpy.uint32 if dtype == numpy.float32 else numpy.uint64
        for form in ("tiled", "flat"):
            monkeypatch.setenv("STENCIL_CUDA_DIRECT", form)
            stencil = make('cuda', mode)
            concrete = stencil.specializer.specialize((grid,) + tables)
            assert concrete.plan.notes['config']['kind'] == form
            actual = stencil(grid, *tables)
            assert (actual.view(view) == expected.view(view)).all(), (form, shape, mode)


def test_out_tensor_is_validated():
    """A caller-supplied out= tensor is handed to the kernel as it stands: wrong shape, dtype,
    device, layout or aliasing must raise instead of writing out of bounds."""
    import torch
    from stencil_code_b200.library.laplacian import LaplacianKernel
    stencil = LaplacianKernel()
    grid = torch.rand((16, 24, 128), device="cuda")
    good = torch.empty_like(grid)
    assert stencil(grid, out=good) is good
    expected = oracle_output(LaplacianKernel(backend='python'), grid.cpu().numpy())
    assert_bit_identical(good.cpu().numpy(), expected)
    for bad in (torch.empty((16, 24, 64), device="cuda"),                       # smaller
                torch.empty((16, 24, 128), device="cuda", dtype=torch.float64),  # wrong dtype
                torch.empty((16, 24, 128)),                                      # host tensor
                torch.empty((16, 24, 256), device="cuda")[:, :, ::2],            # not contiguous
                grid,                                                            # in place
                numpy.empty((16, 24, 128), numpy.float32)):                      # not a tensor
        with pytest.raises(StencilException, match="out="):
            stencil(grid, out=bad)
    overlapping = torch.empty(2 * grid.numel(), device="cuda")
    source = overlapping[:grid.numel()].view(grid.shape)
    with pytest.raises(StencilException, match="overlap"):
        stencil(source, out=overlapping[128:128 + grid.numel()].view(grid.shape))


def test_extreme_magnitudes_need_strict_mode(monkeypatch):
    """The exact float32 demotion of (+-2^k) * x holds while the product neither overflows nor
    becomes subnormal.  Ordinary magnitudes: demoted == strict == oracle.
    Near FLT_MAX / FLT_MIN only STENCIL_CUDA_STRICT=1 (double arithmetic kept) matches the
    reference's C bit for bit -- the documented precondition, pinned here."""
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil       # weights 2, 4, 8, 16
    from stencil_code_b200.library.jacobi_3d import Jacobi3D                         # weight 0.25
    rng = numpy.random.RandomState(3)
    signs = numpy.where(rng.random_sample((24, 128)) < 0.5, -1.0, 1.0)
    huge = (signs * rng.uniform(0.3, 0.9, (24, 128)) * 3.0e38 / 16).astype(numpy.float32)   # 16 * x stays finite
    tiny = (rng.uniform(1.0, 8.0, (8, 16, 128)) * 1.2e-38).astype(numpy.float32)             # 0.25 * x is subnormal
    monkeypatch.setenv("STENCIL_CUDA_STRICT", "1")
    assert_bit_identical(DiagnosticStencil(backend='cuda')(huge),
                         oracle_output(DiagnosticStencil(backend='python'), huge))
    assert_bit_identical(Jacobi3D(1.0, 0.25, backend='cuda')(tiny),
                         oracle_output(Jacobi3D(1.0, 0.25, backend='python'), tiny))
    monkeypatch.delenv("STENCIL_CUDA_STRICT")
    # within the precondition the demoted arithmetic is exact: moderately large and small values
    for scale in (1.0e30, 1.0e-30):
        grid = (rng.uniform(0.5, 2.0, (8, 16, 128)) * scale).astype(numpy.float32)
        assert_bit_identical(Jacobi3D(1.0, 0.25, backend='cuda')(grid),
                             oracle_output(Jacobi3D(1.0, 0.25, backend='python'), grid))


def test_apply_host_on_one_gpu_matches_oracle():
    """SlabStencil.apply_host (pipelined host -> sweep -> host, the multi-GPU e2e path) at world 1,
    followed by iterate(): oracle applied once, then three more times."""
    import stencil_code_b200
    from stencil_code_b200.backend.cuda.multi_gpu import SlabStencil
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    shape = (96, 256, 512)
    window = stencil_code_b200.pinned_empty(shape)
    window[...] = numpy.random.RandomState(5).random_sample(shape).astype(numpy.float32)
    python_side = Jacobi3D(1.0, 0.25, backend='python')
    with SlabStencil(Jacobi3D(1.0, 0.25), shape, world=1, rank=0) as slab:
        assert slab.host_window() == (0, shape[0])
        once = oracle_output(python_side, numpy.array(window), variant='c_parallel')
        assert_bit_identical(numpy.array(slab.apply_host(window)), once)
        slab.iterate(3)
        expected = once
        for _ in range(3):
            expected = oracle_output(python_side, expected, variant='c_parallel')
        assert_bit_identical(numpy.array(slab.result()), expected)


ROW_FUSED_SHAPES = [(40, 256), (37, 388), (130, 1024), (19, 2052), (64, 4096)]


@pytest.mark.parametrize("mode", ["clamp", "zero", "copy"])
@pytest.mark.parametrize("name", ["jacobi2d", "simple", "odd_weights", "conv5", "diagnostic"])
def test_rank2_temporal_blocking_equals_separate_sweeps(name, mode, monkeypatch):
    """iterate() on rank-2 grids fuses pairs of sweeps in registers (codegen_temporal2d): the
    result must be bit-identical to sweeping n times -- checked against the CPU oracle applied n
    times, on widths with one and several CTAs per row, full and partial sub-tiles."""
    from cases import CONV5, OddWeights
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
    from stencil_code_b200.library.jacobi_stencil import Jacobi
    from stencil_code_b200.library.simple import SimpleKernel
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")      # small grids: not "worthwhile"
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    make = {"jacobi2d": lambda backend: Jacobi(backend=backend, boundary_handling=mode),
            "simple": lambda backend: SimpleKernel(backend=backend, boundary_handling=mode),
            "odd_weights": lambda backend: OddWeights(backend=backend, boundary_handling=mode),
            "conv5": lambda backend: ConvolutionFilter(CONV5 / 64.0, backend=backend, boundary_handling=mode),
            "diagnostic": lambda backend: DiagnosticStencil(backend=backend, boundary_handling=mode)}[name]
    for shape in ROW_FUSED_SHAPES:
        grid = (numpy.random.RandomState(29).random_sample(shape) * 0.01).astype(numpy.float32)
        stencil, python_side = make('cuda'), make('python')
        hidden = (stencil.coefficients,) if name == "conv5" else ()
        concrete = stencil.specializer.specialize((grid,) + hidden)
        fused = concrete.temporal()
        assert fused is not None and fused[0].function == "stencil_temporal2_rows", "rank-2 fused kernel expected"
        expected = grid
        for sweeps in (1, 2, 3, 4, 5):
            expected = oracle_output(python_side, expected)
            if sweeps in (2, 3, 5):
                assert_bit_identical(stencil.iterate(grid, sweeps), expected)


def test_rank2_pairs_fuse_into_one_launch(monkeypatch):
    """fuse(a, b) of two different rank-2 stencils: one launch, same bits as b(a(grid))."""
    from cases import OddWeights
    from stencil_code_b200 import fuse
    from stencil_code_b200.library.jacobi_stencil import Jacobi
    from stencil_code_b200.library.simple import SimpleKernel
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    grid = (numpy.random.RandomState(31).random_sample((70, 1028)) * 0.01).astype(numpy.float32)
    for mode in ("clamp", "zero", "copy"):
        first, second = SimpleKernel(boundary_handling=mode), OddWeights(boundary_handling=mode)
        pipeline = fuse(first, second)
        result = pipeline(grid)
        assert pipeline.last_schedule == ['pair']
        expected = oracle_output(OddWeights(backend='python', boundary_handling=mode),
                                 oracle_output(SimpleKernel(backend='python', boundary_handling=mode), grid))
        assert_bit_identical(result, expected)
        three = fuse(first, second, Jacobi(boundary_handling=mode))
        expected = oracle_output(Jacobi(backend='python', boundary_handling=mode), expected)
        assert_bit_identical(three(grid), expected)


@pytest.mark.parametrize("mode", ["clamp", "zero", "copy"])
def test_float64_rank2_grids_on_the_row_pipeline(mode, monkeypatch):
    """float64 rank-2 grids (the reference's known-answer tests use float64,
    test/test_boundary_handling.py:56) stream through TMA like float32 ones: two doubles per
    thread, single sweeps (stencil_rows) and fused pairs (stencil_temporal2_rows), same bits as
    the oracle."""
    from cases import CONV5, OddWeights
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
    from stencil_code_b200.library.jacobi_stencil import Jacobi
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")
    monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
    makes = {"diagnostic": lambda backend: DiagnosticStencil(backend=backend, boundary_handling=mode),
             "jacobi2d": lambda backend: Jacobi(backend=backend, boundary_handling=mode),
             "odd_weights": lambda backend: OddWeights(backend=backend, boundary_handling=mode),
             "conv5": lambda backend: ConvolutionFilter(CONV5 / 64.0, backend=backend, boundary_handling=mode)}
    for name, make in makes.items():
        for shape in ((40, 256), (37, 390), (130, 1024), (19, 2050)):
            grid = numpy.random.RandomState(43).random_sample(shape) * 0.01
            assert grid.dtype == numpy.float64
            stencil, python_side = make('cuda'), make('python')
            hidden = (stencil.coefficients,) if name == "conv5" else ()
            concrete = stencil.specializer.specialize((grid,) + hidden)
            assert concrete.plan.function == "stencil_rows", (name, shape, concrete.plan.function)
            expected = oracle_output(python_side, grid)
            actual = stencil(grid)
            assert actual.dtype == numpy.float64
            assert_bit_identical(actual, expected)
            fused = concrete.temporal()
            assert fused is not None and fused[0].function == "stencil_temporal2_rows"
            for sweeps in (2, 3):
                expected = oracle_output(python_side, expected)
                assert_bit_identical(stencil.iterate(grid, sweeps), expected)
    # the planner takes the row pipeline for large float64 rank-2 grids on its own
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY")
    big = numpy.random.RandomState(44).random_sample((1024, 2048))
    stencil = makes["diagnostic"]('cuda')
    assert stencil.specializer.specialize((big,)).plan.function == "stencil_rows"
    assert_bit_identical(stencil(big), oracle_output(makes["diagnostic"]('python'), big, variant='c_parallel'))


@pytest.mark.parametrize("dtype", [numpy.float32, numpy.float64])
def test_rank1_grids_on_the_vector_kernel(dtype):
    """Large rank-1 grids (reference test/test_boundary_handling.py:145-151 is the 24-point case):
    a thread owns 16 bytes of consecutive points; edge groups, unaligned views and tails run the
    one-point code.  All modes, lengths that are and are not multiples of the vector, a torch view
    that starts 4 bytes into an allocation."""
    import torch
    from cases import Stencil1d
    for mode in ("clamp", "zero", "copy"):
        for length in (4096, 5000, 65537, 1 << 20):
            grid = (numpy.random.RandomState(47).random_sample((length,)) * 1000).astype(dtype)
            stencil = Stencil1d(backend='cuda', boundary_handling=mode)
            concrete = stencil.specializer.specialize((grid,))
            assert concrete.plan.notes['config']['kind'] == 'vector1d'
            expected = oracle_output(Stencil1d(backend='python', boundary_handling=mode), grid)
            assert_bit_identical(stencil(grid), expected)
        whole = torch.from_numpy(numpy.concatenate([numpy.zeros(1, dtype), grid])).cuda()
        view = whole[1:]                                    # contiguous, but not 16-byte aligned
        assert_bit_identical(stencil(view).cpu().numpy(), expected)
