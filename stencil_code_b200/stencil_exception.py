"""StencilException -- the one exception type the stencil API raises.

Mirrors reference ``stencil_code/stencil_exception.py:4-7``.
"""


class StencilException(Exception):
    """Raised for every user-visible error of the stencil DSL and its CUDA backend."""
