#!/bin/bash
mkdir -p gpurun_out
MIG_CONV_SCHED=dynamic timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-sampling --no-hbm --no-cpu-baseline 2> gpurun_out/g8b.err | tee gpurun_out/g8b.log | tail -1 | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-sampling --no-hbm --no-cpu-baseline 2> gpurun_out/g8c.err | tee gpurun_out/g8c.log | tail -1 | cut -c1-200
python - <<PY
import json
for f in ("g8b","g8c"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
    print(f, round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step", {k:round(v["tflops"]) for k,v in d["roofline"]["detail"].items()}, "eager", round(d["roofline"]["eager_ms_per_step"],2))
PY
