"""Top SASS instructions by warp-stall samples from `ncu -i rep --page source --csv` (one block per kernel).
    python scripts/ncu_hot.py file.csv [kernel_index] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[want]
hdr = b["rows"][0]
si, src = hdr.index("# Samples"), hdr.index("Source")
ie = hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in b["rows"][1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print(f"# kernel {want}/{len(blocks)}: {b['name'][:80]}  total samples {tot}, {len(data)} SASS instructions")
agg = {}
for r in data:
    for i in stall:
        if r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("# stall totals:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = {id(r): n for n, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[si]))[:top]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    print(f"{100 * int(r[si]) / tot:5.1f}% #{idx[id(r)]:5d} exec {r[ie]:>9s}  {r[src].strip()[:70]:70s} {st}")
