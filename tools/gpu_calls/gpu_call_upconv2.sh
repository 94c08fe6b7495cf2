#!/bin/bash
# folded upsample conv, second pass: direct-scatter forward, and A/B of direct / class-buffer forward / fold everywhere
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -x -m gpu -k "upconv or upsample or folded" > gpurun_out/u2_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/u2_pytest.log | cut -c1-300
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-hbm --sample-steps 100 > gpurun_out/u2_bench_$label.log 2> gpurun_out/u2_bench_$label.err; echo "bench $label rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/u2_bench_$label.log") if l.startswith("{")][-1])
r=d["roofline"]["detail"]
print("$label", round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step; conv frac", round(d["roofline"]["frac"],3), "| fwd/dgrad/wgrad ms per step", [round(r[k]["seconds"]/3*1e3,2) for k in ("fwd","dgrad","wgrad")], "| sampling", round(d["sampling"]["ms_per_reverse_step"],2), "ms/step")
PY
}
run direct MIG_UPCONV=auto
run classbuf MIG_UPCONV=auto MIG_UPCONV_DIRECT=0
run always MIG_UPCONV=always
run never MIG_UPCONV=never
