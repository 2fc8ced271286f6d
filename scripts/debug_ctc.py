import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ocrs_models_b200 import CTCLoss
import torch.nn.functional as F
T, N, C, S = 201, 64, 97, 40
g = torch.Generator().manual_seed(T * 1000 + N)
lp = torch.log_softmax(torch.randn(T, N, C, generator=g) * 2, dim=2)
tgt = torch.randint(1, C, (N, S), generator=g, dtype=torch.int32)
tgt[:, 1::3] = tgt[:, 0::3][:, : tgt[:, 1::3].shape[1]]
tl = torch.full((N,), S); il = torch.full((N,), T - 1)
lpd = lp.cuda().requires_grad_(True)
loss = CTCLoss()(lpd, tgt.cuda(), il, tl); loss.backward()
lpc = lp.double().requires_grad_(True)
lr = F.ctc_loss(lpc, tgt, il, tl); lr.backward()
d = (lpd.grad.cpu().double() - lpc.grad).abs()
print("loss", loss.item(), lr.item(), "max err", d.max().item(), "gmax", lpc.grad.abs().max().item())
print("err by class (top5):", torch.topk(d.amax(dim=(0, 1)), 5))
print("err by t (top5):", torch.topk(d.amax(dim=(1, 2)), 5))
print("err by n (top5):", torch.topk(d.amax(dim=(0, 2)), 5))
print("row sums ours max:", lpd.grad.sum(2).abs().max().item())
