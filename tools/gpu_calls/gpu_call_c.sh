#!/bin/bash
# GPU call C (2 GPUs): GPU suite, dp_check (fp32 replicated + bf16 sharded), 2-GPU bench sharded vs replicated, 1-GPU bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -6 gpurun_out/c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dp_check.py > gpurun_out/c_dpcheck.log 2>&1; echo "dp_check rc=$?"; grep -E "^dp2|Error|error" gpurun_out/c_dpcheck.log | head
timeout 400 $TR bench.py --gpus 2 --steps 15 --warmup 4 --no-hbm --sample-steps 40 > gpurun_out/c_bench2_shard.log 2> gpurun_out/c_bench2_shard.err; echo "bench2 shard rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 15 --warmup 4 --no-hbm --no-sampling --no-shard > gpurun_out/c_bench2_repl.log 2> gpurun_out/c_bench2_repl.err; echo "bench2 repl rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sample-steps 40 > gpurun_out/c_bench1.log 2> gpurun_out/c_bench1.err; echo "bench1 rc=$?"
python tools/step_profile.py > gpurun_out/c_stepprof.log 2>&1
for f in c_bench2_shard c_bench2_repl c_bench1; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.log") if l.startswith("{")][-1])
    print("$f", round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step sharded=", d.get("optimizer_sharded"), "conv frac", round(d["roofline"]["frac"],3), "sampling", (d.get("sampling") or {}).get("value"))
except Exception as e:
    print("$f", "no line", e)
PY
done
tail -c 600 gpurun_out/c_bench2_shard.err
