# ncu captures for profiles/ (round 2).  Each ncu run follows a plain run of the same command line.
# Run on the GPU box:  bash tools/profile.sh ; summaries are made here with tools/ncu_summary.py
set -u
O=gpurun_out
B="python bench.py --quick --no-cpu --steps 20 --warmup 5"
FULL="ncu --set full --clock-control none --import-source on -f"
# 1. the bench command: plain run, then the launch list (per-launch durations, cold-cache and serialised)
$B > $O/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $O/r02_bench_launches.csv $B > $O/r02_ncu_bench.log 2>&1
# 2. the dominant kernel of the bench command: the fused two-sweep kernel on the full 2048-plane grid
$B > /dev/null 2>&1 && $FULL -k regex:stencil_temporal2 -s 4 -c 1 -o $O/r02_temporal $B > $O/r02_ncu_temporal.log 2>&1
# 3. the other BASELINE configs, one capture each (plain run first)
run() {  # name, kernel regex, skip, command...
    name=$1; regex=$2; skip=$3; shift 3
    "$@" > $O/r02_plain_$name.log 2>&1 && $FULL -k regex:$regex -s $skip -c 1 -o $O/r02_$name "$@" > $O/r02_ncu_$name.log 2>&1
}
run conv5_clamp stencil_rows 3 python tools/probe.py conv5 clamp 5
run conv5_copy stencil_rows 3 python tools/probe.py conv5 copy 5
run bilateral stencil_stream 2 python tools/probe.py bilateral - 3
run laplacian512 stencil_stream 3 python tools/probe.py laplacian512 - 5
run jacobi2d_fused stencil_temporal2_rows 2 python tools/probe_iter2d.py jacobi2d 16384x16384 8
STENCIL_CUDA_TEMPORAL=0 run jacobi3d_single stencil_stream 3 python tools/probe_iter.py 2048x1024x1024 6
ls -la $O/r02_*.ncu-rep
