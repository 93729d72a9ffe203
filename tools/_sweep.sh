timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 tests/multi_gpu_check.py > gpurun_out/r2_mg_check8.log 2>&1; grep -E "MISMATCH|MULTI" gpurun_out/r2_mg_check8.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29513 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -c 300 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n8.json",):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["value"], d["ms_per_step"], d.get("parity"), d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["raw_copy_floor_ms"], d["e2e"]["job_100_sweeps"]["value"], (d.get("one_gpu_same_workload") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
