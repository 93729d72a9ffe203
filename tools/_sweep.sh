timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_benchmarks.py -m gpu -x -q -k "float64 or rank2 or mallard or reference_image or demo" > gpurun_out/r2_t7.log 2>&1; tail -n 6 gpurun_out/r2_t7.log
timeout 300 python tools/probe_f64.py > gpurun_out/r2_probe_f64.log 2>&1; cat gpurun_out/r2_probe_f64.log | tail -12
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; tail -4 gpurun_out/r2_bench_n1b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1b.json").read().strip().split("\n")[-1])
w=d.pop("workloads",{})
print(d["value"], d["ms_per_step"], d["warmup"], d["roofline"]["frac"], d["roofline"]["dram_frac"], d["e2e"]["value"], d["clocks"], d.get("parity"), d["cpu_baseline"]["value"])
for k,v in w.items(): print(k, v.get("gstencil_s"), v.get("frac_of_hbm_peak"), v.get("strategy"), v.get("instruction_roofline",{}).get("frac"))
PY
