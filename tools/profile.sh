# ncu captures for profiles/ (each ncu run follows a plain run of the same command line)
set -u
B="python bench.py --quick --no-cpu --steps 5 --warmup 3"
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_stream -s 4 -c 1 -o gpurun_out/prof_laplacian512_r1c $B > gpurun_out/ncu_b.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench_r1c.csv $B > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*r1c*
