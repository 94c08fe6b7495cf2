#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/f_pytest.log | cut -c1-220
python tools/ae_bench.py > gpurun_out/f_ae.log 2>&1; grep "AE train" gpurun_out/f_ae.log
python tools/sample_profile.py > gpurun_out/f_sampleprof.log 2>&1; head -14 gpurun_out/f_sampleprof.log | cut -c1-150
