#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_custom_ops_gpu.py -m gpu -q -x -k "flash or sdpa or attention or sampling or single_timestep or golden" > gpurun_out/qkv_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/qkv_pytest.log | cut -c1-200
python tools/sample_profile.py A 2>&1 | grep -v Warn | head -12 | cut -c1-150
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm --sample-steps 50 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['sampling']; print(round(d['value'],1),'samples/s; sampling', round(s['ms_per_reverse_step'],3),'ms/step', round(s['value'],2),'vol/min')"
