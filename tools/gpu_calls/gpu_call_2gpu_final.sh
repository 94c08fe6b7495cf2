#!/bin/bash
# 2 GPUs at the head of the round: the driver's launch line for N = 2 (sharded optimiser, CUDA-graph step, folded
# upsample convs), the replica check, and the world-2 sampling block
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-hbm --sample-steps 100"
timeout 600 $RUN 2> gpurun_out/g2_bench.err | tee gpurun_out/g2_bench.log | tail -1 | cut -c1-400; echo "rc=${PIPESTATUS[0]}"
tail -c 500 gpurun_out/g2_bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/g2_dp_check.log 2>&1; echo "dp_check rc=$?"; tail -8 gpurun_out/g2_dp_check.log | cut -c1-250
