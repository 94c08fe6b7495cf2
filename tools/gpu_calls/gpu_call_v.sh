#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "conv" > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/v_pytest.log | cut -c1-220
python tools/narrow_bench.py 10 2>&1 | tee gpurun_out/v_narrow.log
python tools/conv_bench.py 5 all 2>&1 | tee gpurun_out/v_convbench.log | tail -9
python tools/sample_profile.py A 2>&1 | head -14 | cut -c1-160
