"""Back ends that lower the stencil semantic model.  ``cuda`` is the only executable one."""
