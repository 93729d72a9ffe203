P="python tools/probe.py"
$P conv5 clamp 10
STENCIL_CUDA_FIX_BRANCH=0 $P conv5 clamp 10
$P laplacian512 clamp 20
STENCIL_CUDA_FIX_BRANCH=0 $P laplacian512 clamp 20
STENCIL_CUDA_FIX_BRANCH=100 $P laplacian512 clamp 20
$P jacobi3d clamp 10
STENCIL_CUDA_FIX_BRANCH=0 $P jacobi3d clamp 10
