#!/bin/bash
# final validation of the round: whole GPU suite, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log | cut -c1-220
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-200
python bench.py > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/f_bench.err
python bench.py --impl reference > gpurun_out/f_bench_ref.log 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/f_bench.log") if l.startswith("{")][-1])
print(round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms/step; e2e", round(d["e2e"]["value"],1), "conv frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
s=d["sampling"]; print("sampling", round(s["value"],2), "vol/min", round(s["ms_per_reverse_step"],2), "ms/step", s.get("roofline",{}).get("frac"))
print({k:(round(v.get("frac",0),3) if isinstance(v,dict) else v) for k,v in d["hbm_roofline"].items()} if "hbm_roofline" in d else None)
print(d["cpu_baseline"])
r=json.loads([l for l in open("gpurun_out/f_bench_ref.log") if l.startswith("{")][-1]); print("ref", r["value"], r["unit"], r["cpu_baseline"])
PY
