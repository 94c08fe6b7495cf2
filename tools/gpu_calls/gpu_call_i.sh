#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/i_pytest.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sample-steps 40 > gpurun_out/i_bench1.log 2> gpurun_out/i_bench1.err; echo "bench1 rc=$?"; tail -c 600 gpurun_out/i_bench1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/i_bench1.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step conv frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
PY
python tools/step_profile.py > gpurun_out/i_stepprof.log 2>&1; head -34 gpurun_out/i_stepprof.log | cut -c1-150
