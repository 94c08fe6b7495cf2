#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 5 --no-cpu-baseline --no-hbm --no-sampling"
for c in 0 4 8 16; do
  if [ $c = 0 ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  echo "NCCL_MAX_CTAS=$c"
  timeout 600 $RUN 2> gpurun_out/z2_$c.err | tee gpurun_out/z2_$c.log | tail -1 | cut -c1-200
done
grep -h "channels\|NVLS\|nChannels" gpurun_out/z2_0.err | head -5
