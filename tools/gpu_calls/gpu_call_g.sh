#!/bin/bash
# 8-GPU A/B: sharded optimiser (default) vs replicated all-reduce
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-hbm --sample-steps 100 > gpurun_out/g_bench8_shard.log 2> gpurun_out/g_bench8_shard.err; echo "bench8 shard rc=$?"
timeout 420 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-hbm --no-sampling --no-shard > gpurun_out/g_bench8_repl.log 2> gpurun_out/g_bench8_repl.err; echo "bench8 repl rc=$?"
for f in g_bench8_shard g_bench8_repl; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.log") if l.startswith("{")][-1])
    print("$f", round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step sharded=", d.get("optimizer_sharded"), "conv frac", round(d["roofline"]["frac"],3), "eager", round(d["roofline"]["eager_ms_per_step"],1), "sampling", (d.get("sampling") or {}).get("value"))
except Exception as e:
    print("$f", "no line", e)
PY
done
tail -c 400 gpurun_out/g_bench8_shard.err
