"""Per-kernel SASS evidence of the built library: instruction counts that show what the kernels are made of.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTMALDG = TMA tensor loads (cp.async.bulk.tensor), SYNCS = mbarrier, VABSDIFF4 / IDP.4A = packed-byte SAD / SSD,
REDUX = warp reductions, LDGSTS = cp.async, LDS/STS = shared memory, HMMA/UTC*MMA would be tensor cores (none: the
path has no dense contraction).  Static counts of the code, not of executed instructions.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "global-motion-estimation_b200", "libgme_b200.so")
WATCH = ["UTMALDG", "SYNCS", "LDGSTS", "VABSDIFF4", "IDP", "REDUX", "LDS", "STS", "ATOMS", "SHFL", "DFMA", "DADD", "DMUL",
         "HMMA", "UTCHMMA", "UTCIMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = kernels.setdefault(re.sub(r"\(.*", "", name).replace("void gme::", ""), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["total"] += 1
        cur[m.group(1)] += 1
print(f"# {os.path.relpath(LIB, ROOT)}: static SASS instruction counts per kernel (sm_100a)")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{w:>9s}" for w in WATCH))
for name, c in kernels.items():
    print(f"{name[:58]:58s} {c['total']:6d} " + " ".join(f"{c[w]:9d}" for w in WATCH))
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print(f"{'ALL':58s} {tot['total']:6d} " + " ".join(f"{tot[w]:9d}" for w in WATCH))
