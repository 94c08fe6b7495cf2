"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph` into per-kernel shares of
one training step. usage: python tools/summarize_launch_list.py launches.csv [steps_to_use] > summary.csv

The list holds warm-up + timed + instrumented steps back to back; a step is delimited by its single `adamw_kernel`
launch. The last `steps_to_use` COMPLETE steps before the end of the list are averaged."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
use = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = []
with open(path, newline="") as f:
    rd = csv.reader(l for l in f if not l.startswith("=="))
    hdr = next(rd)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[ki]).strip()
        rows.append((name, us))
ends = [i for i, (n, _) in enumerate(rows) if "adamw_kernel" in n and "strided" not in n]
if len(ends) < use + 1:
    sys.exit(f"only {len(ends)} optimiser steps in the list")
lo, hi = ends[-use - 1] + 1, ends[-1] + 1
agg = defaultdict(lambda: [0, 0.0])
for n, us in rows[lo:hi]:
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {len(rows)} launches in the list; rows {lo}..{hi - 1} = {use} whole steps, {(hi - lo) / use:.0f} kernel launches and "
      f"{tot / use / 1e3:.2f} ms of (cold-cache, serialised) kernel time per step. Compare SHARES, not absolutes.")
print("kernel,launches_per_step,ms_per_step,share_pct,avg_us")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{n}\",{c / use:.1f},{us / use / 1e3:.3f},{100 * us / tot:.2f},{us / c:.2f}")
