T="timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
{
$T --master-port 29521 tools/probe_slab.py 100
STENCIL_CUDA_PDL=0 $T --master-port 29522 tools/probe_slab.py 100
STENCIL_CUDA_TEMPORAL_CHUNKS=2 $T --master-port 29523 tools/probe_slab.py 100
STENCIL_CUDA_TEMPORAL_CHUNKS=3 $T --master-port 29524 tools/probe_slab.py 100
STENCIL_CUDA_TEMPORAL_CHUNKS=6 $T --master-port 29525 tools/probe_slab.py 100
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=2,NCW=16,prefetch=4" $T --master-port 29526 tools/probe_slab.py 100
STENCIL_CUDA_STEP_KERNEL=0 $T --master-port 29527 tools/probe_slab.py 100
} 2>&1 | grep "C5 on" | cut -c1-200
