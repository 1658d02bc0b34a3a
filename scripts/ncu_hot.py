"""Developer tool: hottest SASS instructions (warp-stall samples) of a kernel from `ncu --page source --csv` output.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python scripts/ncu_hot.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    why = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]] or 0) > 0}
    why = dict(sorted(why.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {why}")
