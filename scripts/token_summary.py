#!/usr/bin/env python
"""ncu launch list of one token step of the bench `value` (scripts/gpu_bench1.sh: `--metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum`) -> profiles/r02_token_summary.json: launches, DRAM bytes, time and share
per grid, stamped with the hash of the CUDA sources (`bench.csrc_sha16`) so that bench.py only reports `roofline.traffic`
from a list taken on the SAME build."""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

src = sys.argv[1]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
per_launch = defaultdict(dict)
for r in rows:
    per_launch[int(r[0])].update({"kernel": r[4].split("(")[0][-60:], "grid": r[8], r[12]: float(r[14].replace(",", ""))})
groups = defaultdict(lambda: defaultdict(float))
for lid, d in per_launch.items():
    g = groups[f"{d['kernel']} grid {d['grid']}"]
    g["launches"] += 1
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        g[k] += d.get(k, 0.0)
tot_t = sum(g["gpu__time_duration.sum"] for g in groups.values())
out = {"csrc_sha16": bench.csrc_sha16(), "launches": len(per_launch),
       "dram_bytes": sum(g["dram__bytes_read.sum"] + g["dram__bytes_write.sum"] for g in groups.values()),
       "gpu_time_ns_serialised_cold": tot_t,
       "note": "ncu serialises launches and defeats PDL: per-launch times are cold and only their SHARES are meaningful",
       "by_grid": {k: dict(v, share_of_step_pct=round(100 * v["gpu__time_duration.sum"] / tot_t, 1)) for k, v in groups.items()}}
print(json.dumps(out, indent=1))
