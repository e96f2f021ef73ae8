#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell-native tile-movement
instructions (UTMALDG = TMA load, UTMASTG = TMA store, SYNCS = mbarrier transactions), the fp64
pipe mix and the L2 cache-control ops, from `cuobjdump -sass extpom_b200/libpomgpu.so`.

    python scripts/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "extpom_b200", "libpomgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = [("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("CCTL(prefetch)", r"\bCCTL"),
       ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
       ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"), ("MUFU.RCP64H", r"MUFU\.RCP64H"),
       ("MUFU.RSQ64H", r"MUFU\.RSQ64H"), ("BAR", r"\bBAR\.")]
arch = set(re.findall(r"arch = (sm_\w+)", sass))
kern = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            kern[cur][name] += 1
dem = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: cubins for {sorted(arch)}; {len(kern)} kernels")
print("# SASS instruction counts per kernel (static); UTMALDG = cp.async.bulk.tensor load, SYNCS = mbarrier, CCTL = prefetch.global.L1/L2")
cols = [n for n, _ in PAT]
print("%-64s " % "kernel" + " ".join("%7s" % c[:7] for c in cols))
tot = collections.Counter()
for (mangled, cnt), name in zip(kern.items(), dem):
    short = re.sub(r"^void pom::", "", name)
    short = re.sub(r"\(.*$", "", short)
    short = short.replace("pom::", "")
    print("%-64s " % short[:64] + " ".join("%7d" % cnt[c] for c in cols))
    tot.update(cnt)
print("%-64s " % "TOTAL" + " ".join("%7d" % tot[c] for c in cols))
