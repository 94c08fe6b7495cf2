#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 5 --no-cpu-baseline --no-hbm --no-sampling"
timeout 600 $RUN 2> gpurun_out/z_dyn.err | tee gpurun_out/z_dyn.log | tail -1 | cut -c1-330
MIG_CONV_SCHED=static timeout 600 $RUN 2> gpurun_out/z_static.err | tee gpurun_out/z_static.log | tail -1 | cut -c1-330
timeout 600 $RUN 2> gpurun_out/z_dyn2.err | tee gpurun_out/z_dyn2.log | tail -1 | cut -c1-330
