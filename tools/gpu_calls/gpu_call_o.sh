#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/o_pytest.log | cut -c1-200
python tools/kernel_probe.py flash 5 2>&1 | tee gpurun_out/o_flash.log | tail -5
python tools/sample_profile.py > gpurun_out/o_sampleprof.log 2>&1; head -12 gpurun_out/o_sampleprof.log | cut -c1-150
