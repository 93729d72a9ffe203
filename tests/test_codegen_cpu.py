"""Code generation checked without a GPU: every case lowers, plans and NVRTC-compiles for sm_100a
under both strategies; the typing / demotion rules and the planner are unit-tested with a fake
device (the reference tests its planner the same way, ``test/test_local_size_computer.py``)."""
import numpy
import pytest

from cases import CASES, CASE_MODE_PAIRS, PAIR_IDS
from stencil_code_b200.backend import point_program as pp
from stencil_code_b200.backend.cuda import runtime, planner, codegen_stream
from stencil_code_b200.python_frontend import PythonToStencilModel
from stencil_code_b200.stencil_kernel import StencilArgConfig
from stencil_code_b200.stencil_exception import StencilException

BIG_SHAPES = {1: (4096,), 2: (96, 256), 3: (24, 40, 128)}
AUTO_STREAM_SHAPES = {2: (2048, 2048), 3: (128, 128, 256)}     # large enough for the planner to stream


def make_plan(stencil, shapes_dtypes, strategy=None, **kwargs):
    specializer = stencil.specializer
    cfg = tuple(StencilArgConfig(s[0], numpy.dtype(d), len(s), tuple(s)) for s, d in shapes_dtypes)
    model = PythonToStencilModel().visit(specializer.original_tree)
    return specializer.backend(None, None, stencil, arg_cfg=cfg, strategy=strategy, **kwargs).visit(model)


def arg_cfg_for(case, shape):
    extra = [(numpy.asarray(t).shape, numpy.asarray(t).dtype) for t in case.extra]
    stencil = case.make('clamp', 'cuda')
    if hasattr(stencil, 'intensity_lut'):
        extra = [(stencil.distance_lut.shape, stencil.distance_lut.dtype),
                 (stencil.intensity_lut.shape, stencil.intensity_lut.dtype)]
    elif hasattr(stencil, 'coefficients'):
        extra = [(stencil.coefficients.shape, stencil.coefficients.dtype)]
    return [(shape, numpy.float32)] + extra


@pytest.mark.parametrize("strategy", ["direct", "stream"])
@pytest.mark.parametrize("case,mode", CASE_MODE_PAIRS, ids=PAIR_IDS)
def test_every_case_compiles_for_sm100a(case, mode, strategy, monkeypatch):
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    stencil = case.make(mode, 'cuda')
    shape = BIG_SHAPES[len(case.shape)] if strategy == "stream" else case.shape
    if strategy == "stream" and len(shape) == 1:
        pytest.skip("rank 1 has no streaming generator")
    plan = make_plan(stencil, arg_cfg_for(case, shape), strategy)
    assert plan.strategy == strategy
    module = runtime.compile_source(plan.source, "cpu_check.cu", plan.options)
    assert module.cubin()[:4] == b"\x7fELF"
    launch = plan.geometry(*plan.domain.sweep_range)
    assert all(g >= 1 for g in launch.grid)


def test_planner_picks_direct_for_small_and_stream_for_large(monkeypatch):
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from stencil_code_b200.library.simple import SimpleKernel
    assert make_plan(SimpleKernel(), [((1024, 1024), numpy.float32)]).strategy == 'direct'
    assert make_plan(SimpleKernel(), [(AUTO_STREAM_SHAPES[2], numpy.float32)]).strategy == 'stream'
    assert make_plan(LaplacianKernel(), [((24, 40, 128), numpy.float32)]).strategy == 'direct'
    assert make_plan(LaplacianKernel(), [(AUTO_STREAM_SHAPES[3], numpy.float32)]).strategy == 'stream'
    assert make_plan(LaplacianKernel(), [((24, 40, 128), numpy.float64)]).strategy == 'direct'


def test_literal_typing_and_exact_demotion():
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from cases import OddWeights
    plan = make_plan(DiagnosticStencil(), [((16, 16), numpy.float32)], 'direct')
    assert plan.program.demoted == 4 and not plan.program.uses_double()       # 2.0,4.0,8.0,16.0
    plan = make_plan(OddWeights(), [((16, 16), numpy.float32)], 'direct')
    assert plan.program.uses_double()                                          # 0.1 is not 2^k
    stores = [s for s in plan.program.statements if isinstance(s, pp.Store)]
    assert all(s.expr.ctype == pp.DOUBLE for s in stores)     # 0.3 and 0.1 are not float32 values
    plan = make_plan(LaplacianKernel(), [((8, 8, 8), numpy.float32)], 'direct')
    first = plan.program.statements[0].expr
    assert isinstance(first, pp.Bin) and first.left.ctype == pp.INT and first.ctype == pp.FLOAT
    assert "(-4)" in plan.source and "--fmad=false" in plan.options


def test_strict_mode_keeps_doubles(monkeypatch):
    from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
    monkeypatch.setenv("STENCIL_CUDA_STRICT", "1")
    plan = make_plan(DiagnosticStencil(), [((16, 16), numpy.float32)], 'direct')
    assert plan.program.demoted == 0 and plan.program.uses_double()
    assert "2.0 *" in plan.source


def test_wrap_is_zero_as_in_the_reference():
    from stencil_code_b200.library.laplacian import LaplacianKernel
    wrap = make_plan(LaplacianKernel(boundary_handling='wrap'), [((8, 8, 8), numpy.float32)], 'direct')
    zero = make_plan(LaplacianKernel(boundary_handling='zero'), [((8, 8, 8), numpy.float32)], 'direct')
    assert wrap.source == zero.source


def test_planner_with_fake_device(monkeypatch):
    from stencil_code_b200.library.laplacian import LaplacianKernel
    small = planner.DeviceProps(sm_count=4, smem_per_sm=64 * 1024, smem_per_block_optin=48 * 1024)
    cfg = [((512, 512, 512), numpy.float32)]
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    plan_small = make_plan(LaplacianKernel(), cfg, device_props=small)
    plan_b200 = make_plan(LaplacianKernel(), cfg)
    assert plan_small.notes['config']['smem_bytes'] <= 48 * 1024
    assert plan_b200.notes['config']['smem_bytes'] > plan_small.notes['config']['smem_bytes']
    launch = plan_b200.geometry(0, 512)
    slots = 148 * plan_b200.notes['config']['blocks_per_sm']
    items = launch.grid[0] * launch.grid[1] * launch.grid[2]
    waves = -(-items // slots)
    assert items / float(waves * slots) > 0.9            # whole waves of resident CTAs
    assert launch.extra[0] * launch.grid[2] >= 512


def test_chunk_planner_examples():
    assert codegen_stream.plan_chunks(512, 64, 296, 3, 16) == 9       # 576 CTAs ~ 1.95 waves
    assert codegen_stream.plan_chunks(16, 1, 296, 3, 16) == 1
    chunks = codegen_stream.plan_chunks(2048, 256, 296, 3, 16)
    items = 256 * chunks
    assert items / float(-(-items // 296) * 296) > 0.97


def test_unaligned_rows_fall_back_to_direct(monkeypatch):
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    from stencil_code_b200.library.simple import SimpleKernel
    plan = make_plan(SimpleKernel(), [((102, 7), numpy.float32)])
    assert plan.strategy == 'direct'
    with pytest.raises(StencilException):
        make_plan(SimpleKernel(), [((102, 7), numpy.float32)], 'stream')


def test_halo_slabs_cover_matches_enumerator():
    from stencil_code_b200.halo_enumerator import HaloEnumerator
    slabs = planner.halo_slabs((6, 7, 8), (1, 2, 1))
    count = sum((a1 - a0) * (b1 - b0) * (c1 - c0) for (a0, a1), (b0, b1), (c0, c1) in slabs)
    assert count == len(list(HaloEnumerator((1, 2, 1), (6, 7, 8)))) == 6 * 7 * 8 - 4 * 3 * 6
