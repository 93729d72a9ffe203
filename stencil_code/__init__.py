"""Compatibility alias: ``stencil_code`` -> ``stencil_code_b200``.

User code written against ucb-sejits/stencil_code imports ``stencil_code.stencil_kernel``,
``stencil_code.neighborhood``, ``stencil_code.library.<kernel>`` ...; this package maps those
module paths onto the B200 implementation so such code runs unmodified (on ``backend='cuda'``).

Importing this package also makes the reference's backend names -- ``backend='c'``, ``'omp'``,
``'ocl'``, ``'opencl'``, which reference user code passes explicitly and
``BetterBilateralFilter`` defaults to -- select ``'cuda'`` (with a one-time warning) instead of
raising; results follow the reference's C backend.
``STENCIL_CUDA_LEGACY_BACKENDS=raise`` restores the strict behaviour.
"""
import importlib
import sys

import stencil_code_b200 as _impl

_MODULES = [
    "stencil_kernel", "neighborhood", "halo_enumerator", "stencil_exception", "stencil_model",
    "python_frontend", "backend", "backend.stencil_backend", "backend.cuda",
    "library", "library.laplacian", "library.jacobi_stencil", "library.diagnostic_stencil",
    "library.two_d_heat", "library.convolution", "library.bilateral_filter",
    "library.better_bilateral_filter", "library.laplacian_27pt", "library.jacobi_3d",
    "library.simple",
]
for _name in _MODULES:
    sys.modules[__name__ + "." + _name] = importlib.import_module("stencil_code_b200." + _name)

_impl.stencil_kernel.LEGACY_BACKENDS_RUN_ON_CUDA[0] = True

Stencil = _impl.Stencil
Neighborhood = _impl.Neighborhood
HaloEnumerator = _impl.HaloEnumerator
StencilException = _impl.StencilException
