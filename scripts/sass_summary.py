"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
STAS (st.async), LDGSTS (cp.async), HMMA (mma.sync), FFMA2 (packed fp32 FMA), ATOMS (shared-memory atomics).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ocrs_models_b200", "csrc", "libocrs_b200.so")
PAT = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "STAS", "LDGSTS", "HMMA", "FFMA2", "FFMA", "ATOMS", "LDS", "STS", "SHFL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).split("(")[0].replace("void ", "")
        kern = name
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        for p in PAT:
            if op == p or (p in ("UTMALDG", "HMMA", "SYNCS", "LDGSTS", "ATOMS", "UTCHMMA", "LDTM", "UTCBAR", "STAS", "BAR") and op.startswith(p)):
                counts[kern][p] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): static instruction counts per kernel")
print(f"{'kernel':58s} " + " ".join(f"{p:>7s}" for p in PAT))
for k, c in counts.items():
    if sum(c.values()):
        print(f"{k[:58]:58s} " + " ".join(f"{c.get(p, 0):7d}" for p in PAT))
