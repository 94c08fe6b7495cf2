#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_custom_ops_gpu.py -m gpu -q -x -k "flash or sdpa or attention" > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r_pytest.log | cut -c1-220
timeout 120 python tools/kernel_probe.py flash 5 2>&1 | tee gpurun_out/r_flash.log | head -1
MIG_FLASH_POLY=0 timeout 120 python tools/kernel_probe.py flash 5 2>&1 | head -1
MIG_FLASH_PAIR=0 timeout 120 python tools/kernel_probe.py flash 5 2>&1 | head -1
