#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_custom_ops_gpu.py -m gpu -q -k "flash or sdpa or attention" > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log | cut -c1-200
python tools/kernel_probe.py flash 5 2>&1 | tee gpurun_out/q_flash.log | tail -4
python tools/kernel_probe.py flashbwd 5 2>&1 | tee gpurun_out/q_flashbwd.log | tail -8
python tools/attn_bench.py 2>&1 | tee gpurun_out/q_attn_bench.log | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hbm --sample-steps 20 2>&1 | tee gpurun_out/q_bench.log | tail -1 | cut -c1-1500
