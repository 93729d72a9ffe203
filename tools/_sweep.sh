T="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; tail -c 200 gpurun_out/r2_bench_n2_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2_final.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["warmup"], d.get("parity"), "e2e", d["e2e"]["value"], d["e2e"]["raw_copy_floor_ms"], (d.get("one_gpu_same_workload") or {}).get("value"), d["clocks"], d["config"]["preroll_sweeps"])
PY
