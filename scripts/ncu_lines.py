"""Developer tool: per-source-line samples with the stall-reason split (from an .ncu-rep source page)."""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
hdr = None; fn = ""; fname = ""
for r in rows:
    if not r: continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or (pat and pat not in fn): continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (fname, ln)
    text.setdefault(key, r[1].strip()[:90])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "(Not Issued)" not in h and i < len(r):
            try: agg[key][h] += int(r[i])
            except ValueError: pass
    try:
        agg[key]["#inst"] += int(r[hdr.index("Instructions Executed")])
    except ValueError:
        pass
tot = sum(sum(v for k, v in c.items() if k.startswith("stall_")) for c in agg.values()) or 1
tin = sum(c["#inst"] for c in agg.values()) or 1
allst = collections.Counter()
for c in agg.values():
    for k, v in c.items():
        if k.startswith("stall_"): allst[k] += v
print("total samples", tot, {k[6:]: round(v / tot * 100, 1) for k, v in allst.most_common(8)})
lines = sorted(agg.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k.startswith("stall_")))[:top]
for key, c in lines:
    s = sum(v for k, v in c.items() if k.startswith("stall_"))
    why = ", ".join(f"{k[6:]}={v / s * 100:.0f}%" for k, v in c.most_common(4) if k.startswith("stall_") and v)
    print(f"{key[0][:16]:16s}:{key[1]:4d} samp={s / tot * 100:5.1f}% inst={c['#inst'] / tin * 100:5.1f}%  [{why}]  {text[key]}")
