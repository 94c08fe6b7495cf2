#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "flash or sdpa" > gpurun_out/d_pytest_attn.log 2>&1; echo "attn pytest rc=$?"; tail -15 gpurun_out/d_pytest_attn.log | cut -c1-220
timeout 300 python tools/attn_bench.py > gpurun_out/d_attn_bench.log 2>&1; echo "attn bench rc=$?"; cat gpurun_out/d_attn_bench.log | tail -8
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/d_pytest.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sample-steps 40 > gpurun_out/d_bench1.log 2> gpurun_out/d_bench1.err; echo "bench1 rc=$?"
python tools/step_profile.py > gpurun_out/d_stepprof.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/d_bench1.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step conv frac", round(d["roofline"]["frac"],3), "sampling", (d.get("sampling") or {}).get("ms_per_reverse_step"))
PY
