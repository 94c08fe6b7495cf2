#!/bin/bash
# round-2 GPU call A: full GPU test suite, default bench, kernel probe, ncu --set full of the bandwidth / halo / flash kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.log 2>&1; echo "bench rc=$?"
python tools/kernel_probe.py all 4 > gpurun_out/a_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/a_probe.log
NCU="ncu --set full --clock-control none --import-source on"
python tools/kernel_probe.py gn 1 > gpurun_out/a_p_gn.log 2>&1 && $NCU -k regex:gn_ -c 28 -o gpurun_out/a_gn python tools/kernel_probe.py gn 1 > gpurun_out/a_ncu_gn.log 2>&1
python tools/kernel_probe.py adamw 1 > gpurun_out/a_p_adamw.log 2>&1 && $NCU -k regex:'adamw|sumsq' -c 4 -o gpurun_out/a_adamw python tools/kernel_probe.py adamw 1 > gpurun_out/a_ncu_adamw.log 2>&1
python tools/kernel_probe.py flash 1 > gpurun_out/a_p_flash.log 2>&1 && $NCU -k regex:flash_fwd -c 3 -o gpurun_out/a_flash python tools/kernel_probe.py flash 1 > gpurun_out/a_ncu_flash.log 2>&1
python tools/kernel_probe.py halo 1 > gpurun_out/a_p_halo.log 2>&1 && $NCU -k regex:halo -c 8 -o gpurun_out/a_halo python tools/kernel_probe.py halo 1 > gpurun_out/a_ncu_halo.log 2>&1
ls -la gpurun_out
