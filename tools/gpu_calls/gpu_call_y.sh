#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -m gpu -q -x -k "conv or unet or autoencoder or single_timestep or graph or train_step" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/y_pytest.log | cut -c1-220
python tools/conv_bench.py 5 all 2>&1 | tee gpurun_out/y_convbench.log | tail -8
MIG_CONV_SCHED=static python tools/conv_bench.py 5 all 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hbm --sample-steps 20 2>&1 | tee gpurun_out/y_bench.log | tail -1 | cut -c1-400
MIG_CONV_SCHED=static python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hbm --no-sampling 2>&1 | tail -1 | cut -c1-400
