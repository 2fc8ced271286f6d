"""Per-tensor gradient error of one detection train step vs the fp64 oracle (debug aid)."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import functional as O  # noqa: E402
from ocrs_models_b200 import DetectionModel, balanced_cross_entropy_loss  # noqa: E402

N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 96, 80)
g = torch.Generator().manual_seed(int(sys.argv[4]) if len(sys.argv) > 4 else 0)
torch.manual_seed(1234)
m = DetectionModel()
batch = {"image": torch.rand(N, 1, H, W, generator=g) - 0.5, "mask": (torch.rand(N, 1, H, W, generator=g) < 0.1).float()}
sd = {k: v.clone() for k, v in m.state_dict().items()}
out64, loss64, g64, _ = O.train_step_grads("det", sd, batch, torch.float64)
g32 = O.train_step_grads("det", sd, batch, torch.float32)[2]
m = m.cuda().train()
y = m(batch["image"].cuda())
loss = balanced_cross_entropy_loss(y, batch["mask"].cuda())
loss.backward()
gn = torch.sqrt(sum((v ** 2).sum() for v in g64.values()))
rows = []
for k, p in m.named_parameters():
    e = float((p.grad.cpu().double() - g64[k]).norm() / gn)
    e32 = float((g32[k].double() - g64[k]).norm() / gn)
    rows.append((e, e32, k, float(g64[k].norm() / gn)))
rows.sort(reverse=True)
tot = sum(r[0] ** 2 for r in rows) ** 0.5
print(f"global {tot:.2e}; fp32 oracle {sum(r[1] ** 2 for r in rows) ** 0.5:.2e}")
for e, e32, k, n in rows[:14]:
    print(f"{e:.2e} (fp32 {e32:.2e}) |g|/|G| {n:.2e}  {k}")
