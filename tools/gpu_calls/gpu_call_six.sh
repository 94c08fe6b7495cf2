#!/bin/bash
mkdir -p gpurun_out
python tools/conv_bench.py 5 six
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"conv_tma|splitk" -c 3 -s 6 -o gpurun_out/six python tools/conv_bench.py 1 six > gpurun_out/six_ncu.log 2>&1
ncu -i gpurun_out/six.ncu-rep --page details > gpurun_out/six_details.txt 2>/dev/null
ncu -i gpurun_out/six.ncu-rep --page source --csv -c 1 > gpurun_out/six_src.csv 2>/dev/null
ncu -i gpurun_out/six.ncu-rep --page raw --csv > gpurun_out/six_raw.csv 2>/dev/null
rm -f gpurun_out/six.ncu-rep
grep -E "conv_tma_kernel|splitk|Duration|Registers Per|Grid Size|Dynamic Shared|L2 Cache Throughput|DRAM Throughput|Executed Ipc Active|highest-utilized" gpurun_out/six_details.txt | cut -c1-170 | head -40
