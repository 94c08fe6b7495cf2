#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/k_pytest.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sample-steps 40 > gpurun_out/k_bench1.log 2> gpurun_out/k_bench1.err; echo "bench1 rc=$?"; tail -c 600 gpurun_out/k_bench1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/k_bench1.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step conv frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
PY
python tools/step_profile.py > gpurun_out/k_stepprof.log 2>&1; head -12 gpurun_out/k_stepprof.log | cut -c1-150
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-hbm --no-sampling"
$BENCH > gpurun_out/k_bench_plain.log 2> gpurun_out/k_bench_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/k_launches.csv $BENCH > gpurun_out/k_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/k_launches.csv
