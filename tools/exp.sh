P="python tools/probe.py"
$P bilateral - 3
STENCIL_CUDA_STREAM="TY=16,BY=2,NCW=8,blocks=3" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=8,BY=1,NCW=8,blocks=4" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=8,BY=1,NCW=8,blocks=6" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=32,BY=4,NCW=8,blocks=1" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=32,BY=2,NCW=16,blocks=2" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=16,BY=1,NCW=16,blocks=3" $P bilateral - 3
STENCIL_CUDA_STREAM="TY=16,BY=4,NCW=4,blocks=4" $P bilateral - 3
STENCIL_CUDA_NO_ROLL=1 STENCIL_CUDA_STREAM="TY=8,BY=1,NCW=8,blocks=3" $P bilateral - 3
