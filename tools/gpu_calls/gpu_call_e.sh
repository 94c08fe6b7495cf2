#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_custom_ops_gpu.py -m gpu -q > gpurun_out/e_pytest_ops.log 2>&1; echo "custom ops pytest rc=$?"; tail -15 gpurun_out/e_pytest_ops.log | cut -c1-220
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/e_pytest.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sample-steps 40 > gpurun_out/e_bench1.log 2> gpurun_out/e_bench1.err; echo "bench1 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/e_bench1.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step conv frac", round(d["roofline"]["frac"],3), "sampling", (d.get("sampling") or {}).get("ms_per_reverse_step"))
PY
python tools/ae_bench.py > gpurun_out/e_ae.log 2>&1; tail -5 gpurun_out/e_ae.log
python tools/sample_profile.py > gpurun_out/e_sampleprof.log 2>&1; head -25 gpurun_out/e_sampleprof.log | cut -c1-150
