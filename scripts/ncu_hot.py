"""Hot spots of an .ncu-rep source page (SASS view): top stall sites and stall-reason totals.
usage: python scripts/ncu_hot.py x.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
allsamp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", allsamp, " instructions executed", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {s:28s} {v:8d} {100*v/allsamp:5.1f}%")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
print("top sites (index, samples, main stall, SASS):")
for i in sorted(order):
    r = data[i]
    main = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f"  {i:5d} {int(r[ix['# Samples']]):6d} exec={r[ix['Instructions Executed']]:>9s} {main:22s} {r[ix['Source']].strip()[:90]}")
