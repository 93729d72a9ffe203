set -x
P="python tools/probe.py laplacian512"
$P clamp 20
STENCIL_CUDA_L2_PROMOTION=256 $P clamp 20
STENCIL_CUDA_L2_PROMOTION=0 $P clamp 20
$P zero 20
STENCIL_CUDA_STREAM="TY=64,BY=4,NCW=16" $P clamp 20
STENCIL_CUDA_STREAM="TY=64,BY=8,NCW=8" $P clamp 20
STENCIL_CUDA_STREAM="TY=32,BY=2,NCW=16" $P clamp 20
STENCIL_CUDA_STREAM="TY=16,BY=4,NCW=4" $P clamp 20
STENCIL_CUDA_STREAM="TY=16,BY=2,NCW=8" $P clamp 20
STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=2" $P clamp 20
STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=4" $P clamp 20
STENCIL_CUDA_STRATEGY=direct $P clamp 20
python tools/probe.py jacobi3d clamp 10
STENCIL_CUDA_STREAM="TY=64,BY=4,NCW=16" python tools/probe.py jacobi3d clamp 10
python tools/probe.py jacobi3d zero 10
