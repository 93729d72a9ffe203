timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_periodic.py tests/test_gpu_fuzz.py -m gpu -x -q -k "rank1 or stencil1d or 1d or fuzz or periodic" > gpurun_out/r2_t8.log 2>&1; tail -n 5 gpurun_out/r2_t8.log
timeout 300 python tools/probe_f64.py 2>&1 | grep "rank 1"
