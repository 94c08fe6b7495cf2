#!/bin/bash
# ncu --set full of the pair flash kernel (source page included), reduced to CSV on the box
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/kernel_probe.py flash 1 > gpurun_out/s_p_flash.log 2>&1 && $NCU -k regex:flash_pair -c 1 -o gpurun_out/s_flash python tools/kernel_probe.py flash 1 > gpurun_out/s_ncu_flash.log 2>&1
ncu -i gpurun_out/s_flash.ncu-rep --page raw --csv > gpurun_out/s_flash_raw.csv 2>/dev/null
ncu -i gpurun_out/s_flash.ncu-rep --page source --csv -c 1 > gpurun_out/s_flash_src.csv 2>/dev/null
ncu -i gpurun_out/s_flash.ncu-rep --page details > gpurun_out/s_flash_details.txt 2>/dev/null
rm -f gpurun_out/s_flash.ncu-rep
ls -la gpurun_out | head -20
