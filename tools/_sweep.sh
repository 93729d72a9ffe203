set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "temporal or iterate or long_iterated or strict or extreme or out_tensor or apply_host" > gpurun_out/r2_t2.log 2>&1; tail -4 gpurun_out/r2_t2.log
python -m pytest tests/test_fusion.py tests/test_gpu_fullsize.py -m gpu -x -q -k "jacobi or fus" > gpurun_out/r2_t3.log 2>&1; tail -4 gpurun_out/r2_t3.log
python -m pytest tests/test_gpu_baseline_oracle.py -m gpu -x -q -k "c5" > gpurun_out/r2_t4.log 2>&1; tail -4 gpurun_out/r2_t4.log
{
python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_NO_FMA=1 python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_SHFL=0 python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="prefetch=2" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="prefetch=1" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=2,NCW=16,prefetch=1" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=2,NCW=16,prefetch=4" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=4,NCW=8,prefetch=1" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="MY=32,BY=4,NCW=8,prefetch=4" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL_TILE="MY=64,BY=4,NCW=16,prefetch=1" python tools/probe_iter.py 2048x1024x1024 40
STENCIL_CUDA_TEMPORAL=0 python tools/probe_iter.py 2048x1024x1024 40
python tools/probe_iter.py 2048x1024x1024 40 copy
python tools/probe_iter.py 256x1024x1024 40
python tools/probe_iter.py 512x512x512 40
} > gpurun_out/r2_probe1.log 2>&1
cat gpurun_out/r2_probe1.log | cut -c1-150
