#!/bin/bash
# data path after the kernel rewrite + AE trainer with the real PerceptualLoss class
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_gpu.py tests/test_trainer_dropin.py -q -x -m gpu > gpurun_out/d2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/d2_pytest.log | cut -c1-250
timeout 300 python tools/data_bench.py > gpurun_out/d2_bench.log 2> gpurun_out/d2_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/d2_bench.log; tail -c 600 gpurun_out/d2_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:patch_ -c 12 -o gpurun_out/d2_ncu python tools/data_bench.py --probe all > gpurun_out/d2_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/d2_ncu.ncu-rep --page raw --csv > gpurun_out/d2_ncu_raw.csv 2>/dev/null; ls -la gpurun_out/d2_ncu* | cut -c1-120
