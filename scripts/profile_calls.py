"""Per-entry-point device time of one training step (CUDA events), grouped by call order."""
import sys, torch
sys.path.insert(0, '.')
import bench
from ocrs_models_b200 import _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "det"
dev = torch.device("cuda:0")
wl = bench.Workload(kind, dev, 0, 1)
for _ in range(3):
    wl.step_resident()
torch.cuda.synchronize()
_lib.PROFILE = {}
order = []
orig = _lib.call
wl.step_resident()
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
rows = []
for name, evs in prof.items():
    for a, b, m in evs:
        rows.append((name, a.elapsed_time(b), m))
tot = sum(r[1] for r in rows)
print(f"total {tot:.3f} ms over {len(rows)} calls")
agg = {}
for n, ms, m in rows:
    agg.setdefault(n, []).append((ms, m))
for n, lst in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    print(f"{n:36s} {sum(x[0] for x in lst):9.3f} ms  calls {len(lst)}")
    if len(lst) <= 40:
        print("      " + " ".join(f"{x[0]:.2f}" + (f"({x[1]/1e9:.2f}G)" if x[1] else "") for x in lst))
