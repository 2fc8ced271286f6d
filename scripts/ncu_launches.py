"""Aggregate an ncu launch list (--csv --log-file, metrics gpu__time_duration.sum [+ dram bytes]) per kernel.

    python scripts/ncu_launches.py gpurun_out/r01/step_rec.csv [--each REGEX]
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size"))
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[0], {"name": r[ki], "grid": r[gi], "block": r[bi]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    return per


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return name.split("(")[0]


def main():
    per = load(sys.argv[1])
    each = re.compile(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[2] == "--each" else None
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[short(d["name"])]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"# {sys.argv[1]}: {len(per)} launches, {tot / 1e6:.3f} ms serialised (cold-cache per-launch times: compare shares)")
    print(f"{'kernel':58s} {'n':>4s} {'time_us':>10s} {'share':>6s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s} {'GB/s':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        bw = (a[2] + a[3]) / a[1] if a[1] else 0.0
        print(f"{k[:58]:58s} {a[0]:4d} {a[1] / 1e3:10.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f} {bw:7.0f}")
    if each:
        print()
        for i, d in per.items():
            if each.search(d["name"]):
                t = d.get("gpu__time_duration.sum", 0.0)
                rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
                print(f"{i:>5s} {short(d['name'])[:44]:44s} grid {d['grid']:16s} {t / 1e3:9.1f} us rd {rd / 1e6:8.1f} MB wr {wr / 1e6:8.1f} MB {((rd + wr) / t if t else 0):6.0f} GB/s")


if __name__ == "__main__":
    main()
