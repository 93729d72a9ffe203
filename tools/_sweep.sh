timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -n 2 gpurun_out/r2_smoke.log
(time timeout 300 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -n 4 gpurun_out/r2_bench_final.err
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
timeout 100 python bench.py --workload laplacian512 --quick --no-cpu --steps 100 > gpurun_out/r2_bench_lap.json 2> gpurun_out/r2_bench_lap.err; tail -c 200 gpurun_out/r2_bench_lap.err
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2_gpu_tests_final.log 2>&1; tail -n 10 gpurun_out/r2_gpu_tests_final.log
