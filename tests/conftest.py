import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


def rel_l2(a, b):
    import torch

    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_train_step_vs_oracle(kind, model, batch, loss_fn, out_tol=1e-4, loss_tol=1e-4, grad_tol=1e-3, with_fp32=True):
    """One training step of our CUDA module vs the fp64 oracle (SURVEY section 4 protocol): outputs, loss, global
    gradient rel-L2, per-tensor gradients with a floor, post-step BN buffers. Prints the achieved errors and the
    fp32 oracle's own error next to them; returns (err_out, err_loss, err_grad)."""
    import torch

    from oracle import functional as O

    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    out64, loss64, g64, nb64 = O.train_step_grads(kind, sd, batch, torch.float64)
    g32 = O.train_step_grads(kind, sd, batch, torch.float32)[2] if with_fp32 else None
    m = model.cuda().train()
    for p in m.parameters():
        p.grad = None
    out = m(batch["image"].cuda())
    loss = loss_fn(out)
    loss.backward()
    torch.cuda.synchronize()
    e_out = rel_l2(out, out64)
    e_loss = abs(loss.item() - loss64.item()) / abs(loss64.item())
    ours = {k: p.grad for k, p in m.named_parameters()}
    gn = torch.sqrt(sum((v.double() ** 2).sum() for v in g64.values()))
    e_g = float(torch.sqrt(sum(((ours[k].cpu().double() - g64[k]) ** 2).sum() for k in g64)) / gn)
    e_g32 = float(torch.sqrt(sum(((g32[k].double() - g64[k]) ** 2).sum() for k in g64)) / gn) if with_fp32 else float("nan")
    print(f"[{kind} {tuple(batch['image'].shape)}] rel err vs fp64 oracle: out {e_out:.2e} loss {e_loss:.2e} "
          f"global grad {e_g:.2e} (fp32 oracle's own: {e_g32:.2e})")
    assert e_out < out_tol and e_loss < loss_tol and e_g < grad_tol, (e_out, e_loss, e_g)
    floor = gn / len(g64) ** 0.5
    bad = []
    for k in g64:
        e = (ours[k].cpu().double() - g64[k]).norm()
        e32 = (g32[k].double() - g64[k]).norm() if with_fp32 else 0.0
        if not e <= max(1e-3 * g64[k].norm(), 1e-3 * floor, 10 * e32):
            bad.append((k, float(e), float(g64[k].norm()), float(e32)))
    assert not bad, bad
    for k, v in nb64.items():
        if v.is_floating_point():
            assert rel_l2(m.state_dict()[k], v) < 1e-4, k
        else:
            assert int(m.state_dict()[k]) == int(v), k
    return e_out, e_loss, e_g
