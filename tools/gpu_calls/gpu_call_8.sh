#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --sample-steps 100 2> gpurun_out/g8.err | tee gpurun_out/g8.log | tail -1 | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/g8.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step", {k:round(v["tflops"]) for k,v in d["roofline"]["detail"].items()}, "sampling", round(d["sampling"]["value"],1))
PY
