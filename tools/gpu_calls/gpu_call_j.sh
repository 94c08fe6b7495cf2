#!/bin/bash
# evidence run: ncu launch list of the bench command, ncu --set full of the round-2 kernels (reports reduced to CSV)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-hbm --no-sampling"
$BENCH > gpurun_out/j_bench_plain.log 2> gpurun_out/j_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 25000 -c 5200 --csv --log-file gpurun_out/j_launches.csv $BENCH > gpurun_out/j_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/j_launches.csv
NCU="ncu --set full --clock-control none --import-source on"
red() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python tools/kernel_probe.py gn 1 > gpurun_out/j_p_gn.log 2>&1 && $NCU -k regex:gt_ -c 12 -o gpurun_out/j_gn python tools/kernel_probe.py gn 1 > gpurun_out/j_ncu_gn.log 2>&1; red j_gn
python tools/kernel_probe.py flashbwd 1 > gpurun_out/j_p_fb.log 2>&1 && $NCU -k regex:'flash_bwd|flash_fwd|attn_delta' -c 10 -o gpurun_out/j_fb python tools/kernel_probe.py flashbwd 1 > gpurun_out/j_ncu_fb.log 2>&1; red j_fb
python tools/kernel_probe.py conv 1 > gpurun_out/j_p_conv.log 2>&1 && $NCU -k regex:'conv_tma_kernel|wgrad_tma_kernel' -c 6 -o gpurun_out/j_conv python tools/kernel_probe.py conv 1 > gpurun_out/j_ncu_conv.log 2>&1; red j_conv
cat gpurun_out/j_p_gn.log gpurun_out/j_p_fb.log gpurun_out/j_p_conv.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/j_smoke.log
du -sh gpurun_out
