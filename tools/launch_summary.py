"""Dev tool: summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.

  python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n]
Per-launch ncu times are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    rows.append((r["Kernel Name"], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, us in rows:
    m = re.match(r"(?:void )?(?:arp::)?([\w:]+)(<[^(]*>)?", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name
    a = agg.setdefault(key, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':90s} {'n':>5s} {'total us':>10s} {'share':>6s} {'avg':>8s} {'min':>8s} {'max':>8s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {a[0]:5d} {a[1]:10.1f} {a[1] / tot:6.3f} {a[1] / a[0]:8.1f} {a[2]:8.1f} {a[3]:8.1f}")
print(f"total {tot:.1f} us over {len(rows)} launches")
