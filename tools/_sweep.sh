{
timeout 120 python tools/probe.py conv5 zero 20
timeout 120 python tools/probe.py conv5 clamp 20
STENCIL_CUDA_ROWS_TILE="NCW=16,BY=1" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=4,BY=2" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=4,BY=1" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=8,BY=1,blocks=3" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=8,BY=1,blocks=3" timeout 120 python tools/probe.py conv5 clamp 20
STENCIL_CUDA_ROWS_TILE="NCW=8,BY=1,blocks=1" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=16,BY=1,blocks=1" timeout 120 python tools/probe.py conv5 zero 20
STENCIL_CUDA_ROWS_TILE="NCW=16,BY=1,blocks=1" timeout 120 python tools/probe.py conv5 clamp 20
STENCIL_CUDA_ROWS_TILE="NCW=16,BY=1" timeout 120 python tools/probe.py conv5 copy 20
timeout 120 python tools/probe_iter.py 2048x1024x1024 40 zero
timeout 120 python tools/probe_iter.py 2048x1024x1024 40 copy
timeout 120 python tools/probe_iter.py 2048x1024x1024 40 clamp
} 2>&1 | grep -E "GStencil|rror" | cut -c1-260 > gpurun_out/r2_probe7.log
cat gpurun_out/r2_probe7.log | cut -c1-170
