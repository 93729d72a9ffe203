"""Ready-made stencils, one module per kernel as in reference ``stencil_code/library/``."""
