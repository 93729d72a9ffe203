"""stencil_code_b200 -- the ucb-sejits/stencil_code DSL on a B200-native CUDA backend.

The public surface is the reference's: ``Stencil`` (``stencil_kernel``), ``Neighborhood``,
``HaloEnumerator``, ``StencilException`` and the ``library`` kernels.  Importing the package does
not load the CUDA shim; the first call of a ``backend='cuda'`` stencil does.
"""
from .stencil_exception import StencilException
from .neighborhood import Neighborhood
from .halo_enumerator import HaloEnumerator
from .stencil_kernel import Stencil, SpecializedStencil, StencilArgConfig, product

__all__ = ['Stencil', 'SpecializedStencil', 'StencilArgConfig', 'StencilException',
           'Neighborhood', 'HaloEnumerator', 'product']
__version__ = '0.1.0'
