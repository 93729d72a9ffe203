# ncu captures for profiles/ (each ncu run follows a plain run of the same command line)
set -u
B="python bench.py --quick --no-cpu --steps 5 --warmup 3"
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_stream -s 4 -c 1 -o gpurun_out/prof_laplacian512_r1b $B > gpurun_out/ncu_b.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench_r1b.csv $B > gpurun_out/ncu_b2.log 2>&1
J="python tools/probe_iter.py 512x1024x1024 6"
$J > gpurun_out/plain_j.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_temporal2 -s 2 -c 1 -o gpurun_out/prof_jacobi_temporal_r1b $J > gpurun_out/ncu_j.log 2>&1
K="python tools/probe.py jacobi3d clamp 4 512x1024x1024"
$K > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_stream -s 3 -c 1 -o gpurun_out/prof_jacobi_single_r1b $K > gpurun_out/ncu_k.log 2>&1
L="python tools/probe.py bilateral - 2"
$L > gpurun_out/plain_l.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_stream -s 3 -c 1 -o gpurun_out/prof_bilateral_r1b $L > gpurun_out/ncu_l.log 2>&1
M="python tools/probe.py conv5 clamp 4"
$M > gpurun_out/plain_m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stencil_stream -s 3 -c 1 -o gpurun_out/prof_conv5_r1b $M > gpurun_out/ncu_m.log 2>&1
ls -la gpurun_out/*.ncu-rep
