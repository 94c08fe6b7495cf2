#!/bin/bash
# folded upsample + conv (csrc/upconv.cu): parity tests, then A/B of the training step and of the sampling step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -q -x -m gpu -k "upconv or upsample or folded or ldm_width or autoencoder_matches" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/u_pytest.log | cut -c1-300
for mode in never auto; do
  MIG_UPCONV=$mode timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-hbm --sample-steps 200 > gpurun_out/u_bench_$mode.log 2> gpurun_out/u_bench_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/u_bench_$mode.log") if l.startswith("{")][-1])
print("$mode", round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step; conv frac", round(d["roofline"]["frac"],3), "final loss", d.get("final_loss"), "| sampling", round(d["sampling"]["ms_per_reverse_step"],2), "ms/step", round(d["sampling"]["value"],2), "vol/min")
PY
done
timeout 200 python tools/ae_bench.py > gpurun_out/u_ae_auto.log 2>&1; grep "AE train" gpurun_out/u_ae_auto.log | cut -c1-200
MIG_UPCONV=never timeout 200 python tools/ae_bench.py > gpurun_out/u_ae_never.log 2>&1; grep "AE train" gpurun_out/u_ae_never.log | cut -c1-200
