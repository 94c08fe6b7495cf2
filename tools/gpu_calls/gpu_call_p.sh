#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "flash or sdpa" > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log | cut -c1-200
python tools/kernel_probe.py flash 5 2>&1 | tee gpurun_out/p_flash.log | tail -5
python tools/kernel_probe.py flash 5 2>&1 | tail -4
