timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rank2" > gpurun_out/r2_t5.log 2>&1; tail -n 15 gpurun_out/r2_t5.log
{
for k in jacobi2d simple conv5; do
timeout 120 python tools/probe_iter2d.py $k 16384x16384 40
STENCIL_CUDA_TEMPORAL=0 timeout 120 python tools/probe_iter2d.py $k 16384x16384 40
done
timeout 120 python tools/probe_iter2d.py jacobi2d 16384x16384 40 zero
timeout 120 python tools/probe_iter2d.py jacobi2d 8192x8192 40
STENCIL_CUDA_TEMPORAL_TILE="NCW=4,BY=1" timeout 120 python tools/probe_iter2d.py jacobi2d 16384x16384 40
STENCIL_CUDA_TEMPORAL_TILE="NCW=8,BY=1" timeout 120 python tools/probe_iter2d.py jacobi2d 16384x16384 40
STENCIL_CUDA_TEMPORAL_TILE="NCW=4,BY=2,U=6,S=2" timeout 120 python tools/probe_iter2d.py jacobi2d 16384x16384 40
STENCIL_CUDA_TEMPORAL_TILE="NCW=4,BY=2,U=6,S=4" timeout 120 python tools/probe_iter2d.py jacobi2d 16384x16384 40
} 2>&1 | grep -E "GStencil|Error|error" | cut -c1-200 > gpurun_out/r2_probe3.log
cat gpurun_out/r2_probe3.log
