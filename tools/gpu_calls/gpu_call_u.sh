#!/bin/bash
mkdir -p gpurun_out
python tools/narrow_bench.py 10 2>&1 | tee gpurun_out/u_narrow.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tma -c 1 -s 2 -o gpurun_out/u_c11 python tools/narrow_bench.py 1 0 > gpurun_out/u_ncu.log 2>&1
ncu -i gpurun_out/u_c11.ncu-rep --page source --csv -c 1 > gpurun_out/u_c11_src.csv 2>/dev/null
ncu -i gpurun_out/u_c11.ncu-rep --page details > gpurun_out/u_c11_details.txt 2>/dev/null
rm -f gpurun_out/u_c11.ncu-rep
ls -la gpurun_out | tail -5
