"""stencil_code_b200 -- the ucb-sejits/stencil_code DSL on a B200-native CUDA backend.

The public surface is the reference's: ``Stencil`` (``stencil_kernel``), ``Neighborhood``,
``HaloEnumerator``, ``StencilException`` and the ``library`` kernels.  Importing the package does
not load the CUDA shim; the first call of a ``backend='cuda'`` stencil does.
"""
from .stencil_exception import StencilException
from .neighborhood import Neighborhood
from .halo_enumerator import HaloEnumerator
from .stencil_kernel import Stencil, SpecializedStencil, StencilArgConfig, product
from .fusion import fuse, FusedStencils



def pinned_empty(shape, dtype='float32', device_index=None):
    """numpy array in page-locked host memory; pass such arrays to a ``backend='cuda'`` stencil to
    make its host<->device copies asynchronous (additive helper, needs a GPU)."""
    from .backend.cuda.runtime import pinned_empty as _pinned_empty
    return _pinned_empty(shape, dtype, device_index)


__all__ = ['Stencil', 'SpecializedStencil', 'StencilArgConfig', 'StencilException',
           'Neighborhood', 'HaloEnumerator', 'product', 'pinned_empty', 'fuse', 'FusedStencils']
__version__ = '0.1.0'
