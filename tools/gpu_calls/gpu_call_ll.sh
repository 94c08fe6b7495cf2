#!/bin/bash
# ncu launch list of the training bench (eager, so that every kernel is a separate launch), reduced on the box
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-hbm --no-sampling"
$BENCH > gpurun_out/ll_bench_plain.log 2> gpurun_out/ll_bench_plain.err &&
timeout 660 ncu --metrics gpu__time_duration.sum --clock-control none -c 6500 --csv --log-file gpurun_out/ll_launches.csv $BENCH > gpurun_out/ll_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/ll_launches.csv
python tools/summarize_launch_list.py gpurun_out/ll_launches.csv 3 > gpurun_out/ll_summary.csv; head -14 gpurun_out/ll_summary.csv
rm -f gpurun_out/ll_launches.csv
