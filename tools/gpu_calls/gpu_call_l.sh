#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "temb or time_emb" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/l_pytest.log | cut -c1-200
python tools/layer_profile.py > gpurun_out/l_layers.log 2>&1; head -64 gpurun_out/l_layers.log | cut -c1-150
python tools/ae_profile.py > gpurun_out/l_aeprof.log 2>&1; head -30 gpurun_out/l_aeprof.log | cut -c1-150
