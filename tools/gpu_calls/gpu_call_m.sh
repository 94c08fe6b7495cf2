#!/bin/bash
mkdir -p gpurun_out
echo "== 2 epilogue groups, fused epilogue inputs"; python tools/conv_bench.py 10 all epi 2>&1 | tee gpurun_out/m_conv_epi_g2.log | tail -9
echo "== 1 epilogue group, fused epilogue inputs"; MIG_CONV_EPI_GROUPS=1 python tools/conv_bench.py 10 all epi 2>&1 | tee gpurun_out/m_conv_epi_g1.log | tail -9
echo "== 2 groups, plain"; python tools/conv_bench.py 10 all 2>&1 | tee gpurun_out/m_conv_plain_g2.log | tail -9
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "conv" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.log | cut -c1-200
MIG_CONV_EPI_GROUPS=1 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "conv_fwd_bwd" > gpurun_out/m_pytest_g1.log 2>&1; echo "pytest g1 rc=$?"; tail -3 gpurun_out/m_pytest_g1.log | cut -c1-200
