"""Pins the CPU oracle (``oracle/``) to the reference.

1. Against fixtures made by the UNMODIFIED reference's ``backend='python'`` path
   (``tests/golden/make_golden.py``): bit-exact on integer-valued inputs, to the reference's own
   tolerance (``assert_array_almost_equal`` 6 decimals, ``test/test_c_end_to_end.py:24-32``, applied
   relatively) on random inputs -- full grids, halo included.
2. Against every known-answer value in the reference's ``test/test_boundary_handling.py``.
3. The product's ``backend='python'`` interpreter reproduces the fixtures bit for bit (same numpy,
   same statements), so the API class itself is pinned too.
"""
import glob
import os

import numpy
import pytest

from cases import CASES, CASE_MODE_PAIRS, oracle_output
from stencil_code_b200.library.diagnostic_stencil import DiagnosticStencil
from stencil_code_b200.library.laplacian import LaplacianKernel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_PAIRS = [(c, m) for c, m in CASE_MODE_PAIRS if c.golden]
GOLDEN_IDS = ["{}-{}".format(c.name, m) for c, m in GOLDEN_PAIRS]
INEXACT = ("better_bilateral", "bilateral_r3")      # LUT products are not exact in float32
# Kernels whose weights come from self.distance(): under 'clamp' the reference's Python path calls
# distance() on the CLAMPED neighbour (stencil_kernel.py:615-625) while its C path folds the
# weight of the un-clamped offset (backend/stencil_backend.py:116-119), so the two reference back
# ends differ in the halo.  The reference's own cross-backend tests compare interiors only
# (test/test_c_end_to_end.py:24-32); the C path is the parity target.
DISTANCE_WEIGHTED = ("laplacian27", "better_bilateral", "bilateral_r3")


def comparable(case, mode, stencil, array):
    if mode == 'clamp' and case.name in DISTANCE_WEIGHTED:
        return array[stencil.interior_points_slice()]
    return array


def load(case, mode):
    return numpy.load(os.path.join(GOLDEN, "{}__{}.npz".format(case.name, mode)))


def close_like_reference(actual, expected):
    scale = max(1.0, float(numpy.abs(expected).max()))
    numpy.testing.assert_array_almost_equal(actual / scale, expected / scale, decimal=6)


def test_every_fixture_has_a_case():
    names = {c.name for c in CASES if c.golden}
    files = [f for f in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(f).startswith("mallard_")]
    assert len(files) == len(GOLDEN_PAIRS)
    assert {os.path.basename(f).split("__")[0] for f in files} == names


@pytest.mark.parametrize("case,mode", GOLDEN_PAIRS, ids=GOLDEN_IDS)
def test_oracle_c_matches_reference_python_backend(case, mode):
    data = load(case, mode)
    stencil = case.make(mode, 'python')
    exact = comparable(case, mode, stencil, oracle_output(stencil, data["exact_in"], *case.extra))
    expected = comparable(case, mode, stencil, data["exact_out"])
    if case.name in INEXACT:
        close_like_reference(exact, expected)
    else:
        numpy.testing.assert_array_equal(exact, expected)
    close_like_reference(
        comparable(case, mode, stencil, oracle_output(stencil, data["random_in"], *case.extra)),
        comparable(case, mode, stencil, data["random_out"]))


@pytest.mark.parametrize("case,mode", GOLDEN_PAIRS, ids=GOLDEN_IDS)
def test_python_backend_reproduces_fixture(case, mode):
    data = load(case, mode)
    stencil = case.make(mode, 'python')
    numpy.testing.assert_array_equal(stencil(data["exact_in"], *case.extra), data["exact_out"])
    numpy.testing.assert_array_equal(stencil(data["random_in"], *case.extra), data["random_out"])


# ---- known-answer values of reference test/test_boundary_handling.py --------------------------
def test_known_answer_clamped_and_zero():           # :48-73
    ones = numpy.ones([8, 8])                       # float64, as in the reference test
    clamp = oracle_output(DiagnosticStencil(backend='python', boundary_handling='clamp'), ones)
    assert clamp.dtype == numpy.float64 and (clamp == 30).all()
    zero = oracle_output(DiagnosticStencil(backend='python', boundary_handling='zero'), ones)
    assert zero[0, 0] == 0 and (zero[1:-1, 1:-1] == 30).all()
    assert zero[0].sum() == zero[-1].sum() == zero[:, 0].sum() == zero[:, -1].sum() == 0


def test_known_answer_copy_c():                     # :95-115
    ones = numpy.ones([5, 5]).astype(numpy.float32)
    copy = oracle_output(DiagnosticStencil(backend='python', boundary_handling='copy'), ones)
    for edge in (copy[0], copy[4], copy[:, 0], copy[:, 4]):
        assert list(edge) == [1.0] * 5
    assert (copy[1:4, 1:4] == 30).all()
    clamp = oracle_output(DiagnosticStencil(backend='python', boundary_handling='clamp'), ones)
    assert (clamp == 30).all()


def test_known_answer_1d_copy():                    # :142-161
    from cases import Stencil1d
    out = oracle_output(Stencil1d(backend='python', boundary_handling='copy'),
                        numpy.ones(8).astype(numpy.float32))
    assert list(out) == [1.0, 6.0, 6.0, 6.0, 6.0, 6.0, 6.0, 1.0]


def test_known_answer_3d_laplacian_copy():          # :163-184
    stencil = LaplacianKernel(backend='python', boundary_handling='copy')
    grid = numpy.ones([8, 8, 8]).astype(numpy.float32)
    out = oracle_output(stencil, grid)
    for point in stencil.interior_points(out):
        assert out[point] == 2.0
    for point in stencil.halo_points(out):
        assert out[point] == 1.0


def test_omp_variant_is_interior_only_and_agrees_inside():
    """reference backend/omp.py never clamps or copies: halo stays 0, interior equals C."""
    stencil = LaplacianKernel(backend='python', boundary_handling='clamp')
    grid = (numpy.random.RandomState(3).random_sample((8, 9, 10)) * 1000).astype(numpy.float32)
    c_out = oracle_output(stencil, grid, variant='c')
    omp_out = oracle_output(stencil, grid, variant='omp')
    inside = stencil.interior_points_slice()
    numpy.testing.assert_array_equal(omp_out[inside], c_out[inside])
    halo = omp_out.copy()
    halo[inside] = 0
    assert not halo.any()
    numpy.testing.assert_array_equal(oracle_output(stencil, grid, variant='c_parallel'), c_out)


def test_oracle_matches_reference_on_the_mallard_crop():
    """The reference's real-image fixture (``demo/mallard.raw``, 64x64 crop) through the unmodified
    reference's ``BetterBilateralFilter()`` (python backend; ``tests/golden/make_golden.py: mallard``):
    the restated C agrees on the interior to the reference tests' 6 decimals (the python path folds
    distance weights per CLAMPED offset in the halo, the C path per un-clamped offset)."""
    from cases import oracle_output
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    data = numpy.load(os.path.join(GOLDEN, "mallard_crop__clamp.npz"))
    stencil = BetterBilateralFilter(backend='python')                    # the reference's defaults: 3 / 70
    assert tuple(stencil.ghost_depth) == (9, 9)
    actual = oracle_output(stencil, data["image_in"])
    interior = stencil.interior_points_slice()
    close_like_reference(actual[interior], data["image_out"][interior])
