#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n_pytest.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/n_bench_default.log 2> gpurun_out/n_bench_default.err; echo "bench rc=$?"; tail -c 400 gpurun_out/n_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/n_bench_ref.log 2> gpurun_out/n_bench_ref.err; echo "bench ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/n_bench_default.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["value"],1), "conv frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], d["clocks"])
print("cpu", d["cpu_baseline"])
s=d["sampling"]; print("sampling", s["value"], s["ms_per_reverse_step"], s["roofline"]["frac"], s["cpu_baseline"])
for k,v in d["hbm_roofline"]["kernels"].items(): print(k, round(v["avg_launch_ms"]*1e3,1),"us", round(v["gbs"]), "GB/s", round(v["frac"],3))
r=json.loads([l for l in open("gpurun_out/n_bench_ref.log") if l.startswith("{")][-1])
print("ref arm", r["value"], r["cpu_baseline"]["kind"], r.get("sampling"))
PY
