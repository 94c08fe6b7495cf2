#!/bin/bash
# round-2 GPU call A: full GPU test suite, default bench, ncu --set full of the bandwidth / halo / flash kernels.
# ncu reports are reduced to CSV on the box (gpurun_out/ may not exceed 64 MiB).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -8 gpurun_out/a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.log 2> gpurun_out/a_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/a_bench.err
NCU="ncu --set full --clock-control none --import-source on"
red() {  # report -> raw csv (+ source csv for the kernels named in $2), then drop the report
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  if [ -n "$2" ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv -k regex:$2 -c 1 > gpurun_out/$1_src.csv 2>/dev/null; fi
  rm -f gpurun_out/$1.ncu-rep
}
python tools/kernel_probe.py gn 1 > gpurun_out/a_p_gn.log 2>&1 && $NCU -k regex:gn_ -c 14 -o gpurun_out/a_gn python tools/kernel_probe.py gn 1 > gpurun_out/a_ncu_gn.log 2>&1; red a_gn gn_bwd_stats
python tools/kernel_probe.py adamw 1 > gpurun_out/a_p_adamw.log 2>&1 && $NCU -k regex:'adamw|sumsq' -c 2 -o gpurun_out/a_adamw python tools/kernel_probe.py adamw 1 > gpurun_out/a_ncu_adamw.log 2>&1; red a_adamw
python tools/kernel_probe.py flash 1 > gpurun_out/a_p_flash.log 2>&1 && $NCU -k regex:flash_fwd -c 3 -o gpurun_out/a_flash python tools/kernel_probe.py flash 1 > gpurun_out/a_ncu_flash.log 2>&1; red a_flash
python tools/kernel_probe.py halo 1 > gpurun_out/a_p_halo.log 2>&1 && $NCU -k regex:halo -c 6 -o gpurun_out/a_halo python tools/kernel_probe.py halo 1 > gpurun_out/a_ncu_halo.log 2>&1; red a_halo wgrad_halo
du -sh gpurun_out; ls -la gpurun_out
