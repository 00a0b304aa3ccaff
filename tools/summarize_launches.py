"""Summarise an ncu launch list (gpu__time_duration.sum csv) per kernel name: total ms, share, launches."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0.0, 0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    tot[name][0] += ns
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total / 1e6:.2f} ms over {sum(v[1] for v in tot.values())} launches")
for name, (ns, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{ns / 1e6:9.3f} ms {ns / total * 100:5.1f}% {n:5d}x  {name[:110]}")
