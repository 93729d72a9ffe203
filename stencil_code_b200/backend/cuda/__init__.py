"""backend='cuda': lowers the stencil semantic model to sm_100a CUDA source, plans the launch
and runs it through the C-ABI shim (``csrc/libstencil_cuda.so``, ``include/stencil_cuda.h``)."""
