#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 5 --no-cpu-baseline --no-hbm --no-sampling"
echo "coalesced gather"; timeout 600 $RUN 2> gpurun_out/z4_a.err | tee gpurun_out/z4_a.log | tail -1 | cut -c1-200
echo "per-bucket gather"; MIG_COALESCE_GATHER=0 timeout 600 $RUN 2> gpurun_out/z4_b.err | tee gpurun_out/z4_b.log | tail -1 | cut -c1-200
echo "coalesced + NCCL_PROTO=Simple"; NCCL_PROTO=Simple timeout 600 $RUN 2> gpurun_out/z4_c.err | tee gpurun_out/z4_c.log | tail -1 | cut -c1-200
echo "coalesced + NCCL_PROTO=LL128,Simple"; NCCL_PROTO=LL128,Simple timeout 600 $RUN 2> gpurun_out/z4_d.err | tee gpurun_out/z4_d.log | tail -1 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py 2>&1 | tail -6 | cut -c1-200
