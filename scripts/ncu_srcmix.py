"""Instruction mix and hottest SASS regions of one kernel launch in an .ncu-rep (source page).

    python scripts/ncu_srcmix.py REPORT KERNEL_REGEX [launch_skip]
"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
seen = set()
for r in rows[hi + 1:]:
    if len(r) < 8 or r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    data.append(r)
ie, iss, isrc = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
def I(x):
    try: return int(x)
    except Exception: return 0
tot, tots = sum(I(r[ie]) for r in data), max(sum(I(r[iss]) for r in data), 1)
print(f"{rows[0][1][:90] if rows[0] else ''}\n{len(data)} SASS lines, {tot} warp instructions executed, {tots} stall samples")
agg = {}
for r in data:
    t = r[isrc].split()
    if not t: continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    a = agg.setdefault(op, [0, 0]); a[0] += I(r[ie]); a[1] += I(r[iss])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {k:10s} {100 * v[0] / tot:5.1f}% of instructions  {100 * v[1] / tots:5.1f}% of stall samples")
top = sorted(data, key=lambda r: -I(r[iss]))[:14]
print("hottest lines by stall samples:")
for r in top:
    print(f"  {100 * I(r[iss]) / tots:5.1f}%  {r[isrc][:80]}")
