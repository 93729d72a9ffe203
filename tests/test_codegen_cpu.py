"""Code generation checked without a GPU: every case lowers, plans and NVRTC-compiles for sm_100a
under both strategies; the typing / demotion rules and the planner are unit-tested with a fake
device (the reference tests its planner the same way, ``test/test_local_size_computer.py``)."""
import numpy
import pytest

from cases import CASE_MODE_PAIRS, PAIR_IDS
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
    assert make_plan(SimpleKernel(), [((512, 512), numpy.float32)]).strategy == 'direct'
    assert make_plan(SimpleKernel(), [((1024, 1024), numpy.float32)]).strategy == 'stream'
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


# ---- large neighborhoods: tiled plane kernels with re-rolled neighbor loops ------------------------
def bilateral_plan(shape=(256, 512), sigma_d=3):
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    stencil = BetterBilateralFilter(sigma_d, 70)
    return make_plan(stencil, [(shape, numpy.float32), (stencil.distance_lut.shape, numpy.float32),
                               (stencil.intensity_lut.shape, numpy.float32)], 'stream')


def test_bilateral_is_tiled_rolled_and_uses_the_fast_table_index():
    plan = bilateral_plan()
    config = plan.notes['config']
    assert config['embedded_plane'] and config['mode'] == 'smem' and config['static_smem'] == 256 * 128
    roll = codegen_stream.find_rolled_run(codegen_stream.embed_as_plane(plan.program))
    assert roll is not None and len(roll['dys']) == 19 and len(roll['dxs']) == 19
    assert roll['stop'] - roll['start'] == 361
    kinds = [p[0] for p in roll['params'] if p is not None]
    assert kinds == ['table']                       # filter_d[(int)distance] -> per-tap parameter
    source = plan.source
    assert "#pragma unroll 1" in source and "for (int tr_ = 0; tr_ < 19; ++tr_)" in source
    assert "__fadd_rz(fabsf(" in source and "8388608.0f" in source      # abs((int)x) as FADD.RZ
    assert "rep2[" in source and "lds_f32(repk2 + (" in source and "<< 7))" in source   # bank-replicated LUT, one shift-add per gather
    assert len(source) < 80 * 1024                                       # 361 taps are NOT unrolled
    assert runtime.compile_source(source, "bilateral.cu", plan.options).cubin()[:4] == b"\x7fELF"


def test_rolled_loop_with_constant_coefficients():
    from stencil_code_b200.library.convolution import ConvolutionFilter
    kernel = numpy.arange(1, 82, dtype=numpy.float64).reshape(9, 9) / 7.0      # non-symmetric doubles
    stencil = ConvolutionFilter(convolution_array=kernel, boundary_handling='clamp')
    plan = make_plan(stencil, [((96, 256), numpy.float32), ((81,), numpy.float64)], 'stream')
    lifted = codegen_stream.embed_as_plane(plan.program) if plan.notes['config']['embedded_plane'] else plan.program
    roll = codegen_stream.find_rolled_run(lifted)
    if plan.notes['config']['mode'] == 'smem':
        assert roll is not None and [p[0] for p in roll['params'] if p is not None] == ['const']
        assert "__constant__ double cpar" in plan.source
    runtime.compile_source(plan.source, "conv9.cu", plan.options)


def test_small_neighborhoods_are_not_rolled():
    from stencil_code_b200.library.laplacian import LaplacianKernel
    plan = make_plan(LaplacianKernel(), [((24, 40, 128), numpy.float32)], 'stream')
    assert codegen_stream.find_rolled_run(plan.program) is None and "tr_" not in plan.source


# ---- temporal blocking -----------------------------------------------------------------------------
def test_temporal_blocking_eligibility_and_compile():
    from stencil_code_b200.backend.cuda import codegen_temporal
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from stencil_code_b200.library.laplacian_27pt import SpecializedLaplacian27
    from stencil_code_b200.library.simple import SimpleKernel
    shape = (64, 96, 512)
    for mode in ('clamp', 'zero', 'copy'):
        for stencil in (Jacobi3D(boundary_handling=mode), LaplacianKernel(boundary_handling=mode)):
            plan = make_plan(stencil, [(shape, numpy.float32)], 'stream')
            ok, why = codegen_temporal.eligible(plan.program, plan.domain)
            assert ok, why
            fused = codegen_temporal.make_plan(plan.program, plan.domain, planner.B200, plan.options)
            assert fused.strategy == 'temporal' and "bar.sync 1" in fused.source
            runtime.compile_source(fused.source, "temporal.cu", fused.options)
            launch = fused.geometry(0, shape[0])
            config = fused.notes['config']
            assert launch.grid[0] * config['OUT_COLS'] >= shape[2] and launch.grid[1] * config['OUT_ROWS'] >= shape[1]
    plan27 = make_plan(SpecializedLaplacian27(), [(shape, numpy.float32), ((4,), numpy.float32)], 'stream')
    # planes that tile badly (256 columns = three 120-column tiles) are left to single sweeps
    config = codegen_temporal.choose_config(plan.program, planner.B200)
    verdicts = {}
    for shape in ((64, 1024, 1024), (64, 256, 256), (64, 512, 512)):
        program = type("P", (), dict(shape=shape))()
        verdicts[shape] = codegen_temporal.worthwhile(program, config)
    assert verdicts == {(64, 1024, 1024): True, (64, 256, 256): False, (64, 512, 512): True}
    assert not codegen_temporal.eligible(plan27.program, plan27.domain)[0]       # corner taps at dz != 0
    plan2d = make_plan(SimpleKernel(), [((2048, 2048), numpy.float32)], 'stream')
    assert not codegen_temporal.eligible(plan2d.program, plan2d.domain)[0]
    slab = make_plan(Jacobi3D(), [(shape, numpy.float32)], 'stream', open_faces=(True, False))
    assert not codegen_temporal.eligible(slab.program, slab.domain)[0]


def test_slab_domain_constants():
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    plan = make_plan(Jacobi3D(), [((34, 64, 256), numpy.float32)], 'stream', open_faces=(True, True))
    domain = plan.domain
    assert domain.sweep_range == (1, 33) and domain.clamp_lo[0] == 0 and domain.clamp_hi[0] == 33
    top = make_plan(Jacobi3D(), [((34, 64, 256), numpy.float32)], 'stream', open_faces=(False, True)).domain
    assert top.clamp_lo[0] == 1 and top.interior_lo[0] == 2 and top.interior_hi[0] == 34
    assert "if (out_peer)" in plan.source
    single = make_plan(Jacobi3D(), [((34, 64, 256), numpy.float32)], 'stream')
    assert "if (out_peer)" not in single.source        # the mirror store exists only in slab kernels


def test_dependent_launch_variant_compiles_for_every_kernel_family(tmp_path):
    """iterate() uses modules compiled with -DSC_PDL: same source, griddepcontrol instructions in
    (SASS: PREEXIT = launch_dependents, ACQBULK = wait); the module for single calls has neither."""
    import shutil
    import subprocess
    from stencil_code_b200.backend.cuda import codegen_temporal
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    from stencil_code_b200.library.simple import SimpleKernel
    plans = [make_plan(Jacobi3D(), [((64, 96, 512), numpy.float32)], 'stream'),
             make_plan(Jacobi3D(), [((64, 96, 512), numpy.float32)], 'direct'),
             make_plan(SimpleKernel(), [((300, 1024), numpy.float32)], 'stream')]
    plans.append(codegen_temporal.make_plan(plans[0].program, plans[0].domain, planner.B200, plans[0].options))
    for plan in plans:
        assert "pdl_wait();" in plan.source
        plain = runtime.compile_source(plan.source, "plain.cu", plan.options).cubin()
        dependent = runtime.compile_source(plan.source, "dependent.cu", list(plan.options) + ["-DSC_PDL"]).cubin()
        assert plain != dependent
        if shutil.which("cuobjdump"):
            path = tmp_path / "k.cubin"
            for data, expected in ((plain, False), (dependent, True)):
                path.write_bytes(data)
                sass = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True).stdout
                assert ("ACQBULK" in sass and "PREEXIT" in sass) == expected     # wait / launch_dependents


def test_host_pipeline_chunk_bounds():
    """Chunks of the pipelined numpy path: cover [lo, hi) exactly, none thinner than four ghost
    depths (a chunk's sweep reads ghost planes of its neighbours only), short first and last."""
    bounds_of = runtime.ConcreteCudaStencil._pipeline_bounds
    for lo, hi, chunks, thin in ((0, 512, 16, 4), (0, 512, 32, 4), (0, 100, 3, 4), (1, 63, 2, 8),
                                 (0, 2048, 64, 4), (0, 16, 4, 4), (0, 7, 1, 4), (3, 40, 5, 36)):
        bounds = bounds_of(lo, hi, chunks, thin)
        sizes = [b - a for a, b in zip(bounds, bounds[1:])]
        assert bounds[0] == lo and bounds[-1] == hi and all(size > 0 for size in sizes)
        assert len(sizes) == 1 or min(sizes) >= min(thin, (hi - lo) // max(1, chunks))
        assert sizes[0] <= max(sizes) and sizes[-1] <= max(sizes)
    assert [b - a for a, b in zip(bounds_of(0, 512, 16, 4), bounds_of(0, 512, 16, 4)[1:])][:3] == [4, 8, 16]


def test_padded_row_plans_address_by_pitch_and_clamp_inside_the_last_quad():
    """iterate() on rows that are no multiple of 16 bytes: the sibling specialisation keeps the
    true extent in the tensor map (TMA zero-fills past it), strides and store addresses use the
    padded pitch, and under 'clamp' the column cascade starts inside the thread's quad."""
    from cases import CONV5
    from stencil_code_b200.backend.cuda import codegen_temporal
    from stencil_code_b200.library.convolution import ConvolutionFilter
    from stencil_code_b200.library.jacobi_3d import Jacobi3D

    class Fake(object):
        def __init__(self, shape, dtype=numpy.float32):
            self.shape, self.dtype, self.ndim = shape, numpy.dtype(dtype), len(shape)
    stencil = Jacobi3D(1.0, 0.25)
    spec = stencil.specializer
    config = spec.args_to_subconfig((Fake((40, 66, 513)),))
    dense = spec.transform(config)
    assert dense.strategy == 'direct' and not dense.domain.is_pitched          # TMA stride rule
    padded = spec.transform(config, row_pitch=516)
    assert padded.strategy == 'stream' and padded.domain.row_pitch == 516 and padded.program.shape[-1] == 513
    (tensor_map,) = padded.tensor_maps
    assert tensor_map.dims == (513, 66, 40) and tensor_map.strides_bytes == (516 * 4, 66 * 516 * 4)
    assert ") * 516 + gx" in padded.source and "if (gx + 1 > 512)" in padded.source
    runtime.compile_source(padded.source, "padded.cu", padded.options)
    fused = codegen_temporal.make_plan(padded.program, padded.domain, planner.B200, padded.options)
    assert fused.tensor_maps[0].strides_bytes == (516 * 4, 66 * 516 * 4) and ") * 516 + gx" in fused.source
    runtime.compile_source(fused.source, "padded_fused.cu", fused.options)
    aligned = spec.transform(spec.args_to_subconfig((Fake((40, 66, 512)),)))
    assert "if (gx + 1 > 511)" not in aligned.source                           # rows end on a quad boundary
    with pytest.raises(ValueError):
        spec.transform(config, row_pitch=512)                                  # pitch shorter than a row

    conv = ConvolutionFilter(CONV5, boundary_handling='clamp')
    config = conv.specializer.args_to_subconfig((Fake((700, 1025)), Fake((25,), numpy.int64)))
    padded = conv.specializer.transform(config, row_pitch=1028)
    assert padded.strategy == 'stream' and padded.tensor_maps[0].strides_bytes == (1028 * 4,)
    runtime.compile_source(padded.source, "padded2d.cu", padded.options)


def test_point_kernel_tiles_share_tap_loads():
    """'direct' on rank-2/3 grids: RZ x RY points per thread, each distinct tap value loaded once;
    short rows, rank 1 and STENCIL_CUDA_DIRECT=flat keep one point per thread."""
    from stencil_code_b200.backend.cuda import codegen_direct
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from cases import Stencil1d
    plan = make_plan(LaplacianKernel(boundary_handling='clamp'), [((64, 96, 256), numpy.float32)], 'direct')
    config = plan.notes['config']
    assert config == {'kind': 'tiled', 'RY': 8, 'RZ': 2}
    loads = [line for line in plan.source.splitlines() if line.strip().startswith("const float v0_")]
    assert len(loads) == 68 and len(set(loads)) == 68                 # 7 taps x 16 points = 112 without sharing
    assert plan.geometry(0, 64).grid == (2, 12, 32) and plan.geometry(10, 13).grid == (2, 12, 2)
    runtime.compile_source(plan.source, "tiled.cu", plan.options)
    double = make_plan(LaplacianKernel(boundary_handling='copy'), [((64, 96, 256), numpy.float64)], 'direct')
    assert double.notes['config'] == {'kind': 'tiled', 'RY': 4, 'RZ': 2}      # doubles count twice
    runtime.compile_source(double.source, "tiled64.cu", double.options)
    assert make_plan(LaplacianKernel(), [((64, 96, 64), numpy.float32)], 'direct').notes['config']['kind'] == 'flat'
    assert make_plan(Stencil1d(), [((1000,), numpy.float32)], 'direct').notes['config']['kind'] == 'flat'
    assert make_plan(Stencil1d(), [((4096,), numpy.float32)], 'direct').notes['config'] == {'kind': 'vector1d', 'V': 4}
    assert codegen_direct.direct_config(plan.program, plan.domain)['kind'] == 'tiled'
