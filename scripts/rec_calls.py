"""Per-call device times of one recognition train step in call order (CUDA events around every C-ABI call)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from ocrs_models_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bench.REC["n"] = n
dev = torch.device("cuda:0")
wl = bench.Workload("rec", dev, 0, 1)
for _ in range(3):
    wl.step_resident()
torch.cuda.synchronize()
order = []
orig = _lib.call


def call(name, *a, meta=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *a)
    e1.record()
    order.append((name, e0, e1, meta))


_lib.call = call
import ocrs_models_b200.rec_engine as de  # noqa: E402
import ocrs_models_b200.losses as lo  # noqa: E402
import ocrs_models_b200.optim as op  # noqa: E402

de.call = lo.call = op.call = call
wl.step_resident()
torch.cuda.synchronize()
tot = 0.0
for name, e0, e1, meta in order:
    ms = e0.elapsed_time(e1)
    tot += ms
    if ms > 0.03:
        gb = f"{meta / 1e9:7.2f} GFLOP {meta / ms / 1e9:6.1f} TFLOP/s" if meta else ""
        print(f"{name.replace('ocrs_', ''):28s} {ms * 1e3:8.0f} us  {gb}")
print(f"total {tot:.2f} ms over {len(order)} calls")
