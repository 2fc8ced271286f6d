import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_rec_gpu import _rec_model, _rec_batch
from oracle import functional as O
from ocrs_models_b200 import CTCLoss
for (N, W, S) in [(3, 96, 8), (5, 64, 4), (5, 96, 4), (4, 64, 4), (5, 64, 8)]:
    m = _rec_model()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    batch = _rec_batch(N, W, S, torch.Generator().manual_seed(0))
    out64, loss64, g64, nb64 = O.train_step_grads("rec", sd, batch, torch.float64)
    m = m.cuda().train()
    lp = m(batch["image"].cuda())
    loss = CTCLoss()(lp, batch["targets"].cuda(), batch["input_lengths"], batch["target_lengths"])
    loss.backward()
    print("case", N, W, S, "lp err", float((lp.cpu().double() - out64).norm() / out64.norm()), "loss", loss.item(), loss64.item())
    for k, p in m.named_parameters():
        e = (p.grad.cpu().double() - g64[k]).norm() / g64[k].norm().clamp_min(1e-12)
        print(f"   {k:32s} rel {float(e):.2e} norm {float(g64[k].norm()):.3e}")
