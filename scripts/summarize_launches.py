"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table for profiles/.
Usage: python scripts/summarize_launches.py launches.csv [first_launch_id] > profiles/x.md"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
k_i, v_i, u_i, id_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r is hdr or not r[id_i].isdigit() or int(r[id_i]) < first:
        continue
    v = float(r[v_i].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[u_i], 1.0)
    tot[r[k_i][:72]] += v
    cnt[r[k_i][:72]] += 1
total = sum(tot.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {k} | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% |")
