#!/bin/bash
# GPU call B: full GPU suite, default bench, kernel breakdown of the step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -12 gpurun_out/b_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.log 2> gpurun_out/b_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/b_bench.err
python tools/step_profile.py > gpurun_out/b_stepprof.log 2>&1; echo "stepprof rc=$?"; head -40 gpurun_out/b_stepprof.log
