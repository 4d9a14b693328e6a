"""Kernel shares of a `ncu --metrics gpu__time_duration.sum --csv` launch list (read here, no GPU).
    python tools/launch_shares.py gpurun_out/launches_bench.csv "<command>" > profiles/rN_launch_shares.json"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("fitv2::", "")
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
out = dict(command=sys.argv[2] if len(sys.argv) > 2 else "", note="per-launch times are cold-cache and serialised; compare SHARES",
           total_us=tot, kernels=[dict(kernel=k, launches=v[0], total_us=round(v[1], 1), us_per_launch=round(v[1] / v[0], 2), share=round(v[1] / tot, 4))
                                  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])])
print(json.dumps(out, indent=1))
