P="python tools/probe.py"
$P laplacian512 clamp 20
STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=2" $P laplacian512 clamp 20
STENCIL_CUDA_STORE_OP=".cs" STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=2" $P laplacian512 clamp 20
STENCIL_CUDA_CLUSTER=4 STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=2" $P laplacian512 clamp 20
STENCIL_CUDA_CLUSTER=2 STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,prefetch=2" $P laplacian512 clamp 20
$P jacobi3d clamp 10
STENCIL_CUDA_STORE_OP=".cs" $P jacobi3d clamp 10
STENCIL_CUDA_CLUSTER=4 $P jacobi3d clamp 10
STENCIL_CUDA_CLUSTER=8 $P jacobi3d clamp 10
STENCIL_CUDA_CLUSTER=8 STENCIL_CUDA_STORE_OP=".cs" $P jacobi3d clamp 10
STENCIL_CUDA_STREAM="TY=64,BY=4,NCW=16,prefetch=2" $P jacobi3d clamp 10
