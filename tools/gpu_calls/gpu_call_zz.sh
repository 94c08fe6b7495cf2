#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "flash_attention" > gpurun_out/zz_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/zz_pytest.log | cut -c1-200
python tools/sample_profile.py A > gpurun_out/zz_sample_profile.log 2>&1; head -16 gpurun_out/zz_sample_profile.log | cut -c1-150
python tools/sample_layer_profile.py A > gpurun_out/zz_sample_layers.log 2>&1; head -5 gpurun_out/zz_sample_layers.log
python tools/narrow_bench.py 10 > gpurun_out/zz_narrow.log 2>&1
python tools/kernel_probe.py flash 5 > gpurun_out/zz_flash.log 2>&1; cat gpurun_out/zz_flash.log | tail -5
