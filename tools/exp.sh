python tools/probe_iter.py 512x1024x1024 20
STENCIL_CUDA_TEMPORAL=0 python tools/probe_iter.py 512x1024x1024 20
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=4,NCW=8" python tools/probe_iter.py 512x1024x1024 20
STENCIL_CUDA_TEMPORAL_TILE="MY=64,BY=4,NCW=16" python tools/probe_iter.py 512x1024x1024 20
python tools/probe_iter.py 512x1024x1024 20 copy
