#!/bin/bash
# register-resident softmax rows + fused dS narrowing: tests, then the training bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_custom_ops_gpu.py -q -x -m gpu -k "softmax or sdpa or attention or flash" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-hbm --no-sampling > gpurun_out/s_bench.log 2> gpurun_out/s_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s_bench.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step; conv frac", round(d["roofline"]["frac"],3), "final loss", d.get("final_loss"))
PY
timeout 300 python tools/step_profile.py > gpurun_out/s_step_profile.log 2>&1; head -40 gpurun_out/s_step_profile.log | cut -c1-150
