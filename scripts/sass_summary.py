"""Developer tool: which Blackwell-native instructions each kernel of libliger_b200.so contains (SASS mnemonics:
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy),
SYNCS = mbarrier, FFMA2 = packed fp32 FMA, REDG = red.global).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "liger_b200", "lib", "libliger_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(set(re.findall(r"Function : (\S+)", sass))), capture_output=True, text=True)
want = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "DFMA", "RED", "ATOMS", "LDGSTS", "HMMA"]
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1).split(".")[0]
        if op in want:
            counts[fn][op] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass; sm_100a)")
print(f"# {'kernel':58s} " + " ".join(f"{w:>7s}" for w in want))
for fn, c in counts.items():
    if not any(c.values()):
        continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("lgr::", "")
    print(f"{dem[:60]:60s} " + " ".join(f"{c.get(w, 0):7d}" for w in want))
# excerpt: the issue loop of the dense B sweep
print("\n# excerpt: MMA issue section of k_dense_ltx (tcgen05.mma -> UTCHMMA with uniform-register descriptors, tcgen05.commit -> UTCBAR)")
one = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3lgr11k_dense_ltxILi32EEEvPKNS_7DenseDsEiii", lib], capture_output=True, text=True).stdout.splitlines()
idx = [i for i, l in enumerate(one) if "UTCHMMA" in l]
if idx:
    for l in one[max(idx[0] - 12, 0):idx[-1] + 8]:
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if m:
            print(f"  /*{m.group(1)}*/ {m.group(2).strip()}")
