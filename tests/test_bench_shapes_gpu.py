"""Parity of the CUDA path at the shapes bench.py times and the shapes the real training scripts produce:
rec 64x800 (T=201) at batch 4 (golden from the unmodified reference) and batch 64, the collate shape W=1024
(T=257, ragged input lengths, train_rec.py:267-272), det at 1x1024x1024 and at the training size 2x800x600
(train_detection.py:22-24; odd crops 75/37, 25/12), and the GRU recurrence alone at T=201/257."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, check_train_step_vs_oracle, rel_l2
from oracle import functional as O

pytestmark = pytest.mark.gpu


def _rec(seed=1234):
    from ocrs_models_b200 import RecognitionModel

    torch.manual_seed(seed)
    return RecognitionModel(O.DEFAULT_ALPHABET)


def test_rec_kat_800_golden_from_the_reference(golden):
    """SURVEY 8c KAT-REC, recorded by oracle/make_golden.py from the unmodified reference: 4x64x800, T=201."""
    from ocrs_models_b200 import CTCLoss

    gold = golden("rec_kat_800")
    m = _rec().cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(4, 1, 64, 800, generator=g) - 0.5
    lp = m(x.cuda())
    assert lp.shape == (201, 4, 97)
    loss = CTCLoss()(lp, torch.from_numpy(gold["targets"]).cuda(), torch.from_numpy(gold["il"]), torch.from_numpy(gold["tl"]))
    loss.backward()
    mean, std = float(lp.mean()), float(lp.std())
    assert abs(mean - gold["lp_stats"][0]) < 1e-4 * abs(gold["lp_stats"][0]) and abs(std - gold["lp_stats"][1]) < 1e-3 * gold["lp_stats"][1]
    assert abs(loss.item() - float(gold["loss"])) < 1e-4 * float(gold["loss"])  # 66.07199860
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())).item()
    assert abs(gn - float(gold["grad_norm"])) < 1e-3 * float(gold["grad_norm"])
    worst = 0.0
    for k, p in m.named_parameters():
        s = gold["grad.stat." + k]  # sum, abs-sum, L2 of the reference's fp32 gradient
        worst = max(worst, abs(p.grad.double().norm().item() - s[2]) / max(s[2], 1e-3 * float(gold["grad_norm"])))
        if "grad.full." + k in gold:
            ref = torch.from_numpy(gold["grad.full." + k])
            assert (p.grad.cpu() - ref).norm() <= 2e-3 * max(ref.norm(), 1e-3 * float(gold["grad_norm"])), k
    print(f"rec_kat_800: loss {loss.item():.6f} vs {float(gold['loss']):.6f}; grad norm {gn:.5f} vs {float(gold['grad_norm']):.5f}; "
          f"worst per-tensor norm rel diff {worst:.2e}")
    assert worst < 5e-3


@pytest.mark.parametrize("N,W,S,ragged", [(4, 800, 40, True), (64, 800, 40, False), (8, 1024, 64, True)])
def test_rec_train_step_at_bench_and_collate_shapes(N, W, S, ragged):
    from ocrs_models_b200 import CTCLoss

    g = torch.Generator().manual_seed(5)
    batch = {"image": torch.rand(N, 1, 64, W, generator=g) - 0.5,
             "targets": torch.randint(1, 97, (N, 64), generator=g, dtype=torch.int32),
             "input_lengths": torch.full((N,), 200, dtype=torch.int64),  # 800-px lines: W // 4, also when padded to 1024
             "target_lengths": torch.full((N,), S, dtype=torch.int64)}
    if ragged:
        batch["input_lengths"] = torch.randint(150, min(W // 4, 256) + 1, (N,), generator=g)
        batch["input_lengths"][0] = W // 4
        batch["target_lengths"] = torch.randint(1, S + 1, (N,), generator=g)
    ctc = CTCLoss()
    check_train_step_vs_oracle("rec", _rec(), batch, lambda lp: ctc(lp, batch["targets"].cuda(), batch["input_lengths"], batch["target_lengths"]),
                               with_fp32=(N <= 8))


@pytest.mark.parametrize("N,H,W", [(1, 1024, 1024), (2, 800, 600)])
def test_det_train_step_at_bench_and_training_sizes(N, H, W):
    from ocrs_models_b200 import DetectionModel, balanced_cross_entropy_loss

    torch.manual_seed(1234)
    m = DetectionModel()
    g = torch.Generator().manual_seed(0)
    batch = {"image": torch.rand(N, 1, H, W, generator=g) - 0.5, "mask": (torch.rand(N, 1, H, W, generator=g) < 0.1).float()}
    check_train_step_vs_oracle("det", m, batch, lambda y: balanced_cross_entropy_loss(y, batch["mask"].cuda()))


@pytest.mark.parametrize("T,N", [(201, 64), (257, 70)])
def test_gru_recurrence_long_sequences(T, N):
    """Drift of the 201/257-step fp32 recurrence and its BPTT vs the fp64 oracle (models.py:245,264-266)."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr
    from ocrs_models_b200.rec_engine import gemm

    st = _lib.stream_ptr(torch.device("cuda:0"))
    g = torch.Generator().manual_seed(T)
    I, Hd = 128, 256
    x = torch.randn(T, N, I, generator=g)
    P = {}
    for sfx in ("", "_reverse"):
        for nm, shp in (("w_ih", (3 * Hd, I)), ("w_hh", (3 * Hd, Hd)), ("b_ih", (3 * Hd,)), ("b_hh", (3 * Hd,))):
            P[nm + sfx] = (torch.rand(shp, generator=g) * 2 - 1) / 16  # nn.GRU init range 1/sqrt(256)
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    x64 = x.double().requires_grad_(True)
    ref = torch.cat([O._gru_direction(x64, P64["w_ih" + s], P64["w_hh" + s], P64["b_ih" + s], P64["b_hh" + s], r)
                     for s, r in (("", False), ("_reverse", True))], dim=2)
    dout = torch.randn(T, N, 2 * Hd, generator=g)
    ref.backward(dout.double())
    D = {k: v.cuda() for k, v in P.items()}
    xd = x.cuda()
    gi = [gemm(xd, I, True, D["w_ih" + s], I, True, T * N, 768, I, st, bias=D["b_ih" + s]) for s in ("", "_reverse")]
    out = torch.empty(T, N, 512, device="cuda")
    gates = torch.empty(T, N, 2, 4, 256, device="cuda")
    call("ocrs_gru_layer_fwd_persist", ptr(gi[0]), ptr(gi[1]), ptr(D["w_hh"]), ptr(D["w_hh_reverse"]), ptr(D["b_hh"]),
         ptr(D["b_hh_reverse"]), ptr(out), ptr(gates), T, N, st)
    e_f = rel_l2(out, ref)
    whhT = [D["w_hh" + s].t().contiguous() for s in ("", "_reverse")]
    dgi = [torch.empty(T * N, 768, device="cuda") for _ in range(2)]
    dgh = [torch.empty(T * N, 768, device="cuda") for _ in range(2)]
    dd = dout.cuda()
    call("ocrs_gru_layer_bwd_persist", ptr(whhT[0]), ptr(whhT[1]), ptr(dd), ptr(out), ptr(gates), ptr(dgi[0]), ptr(dgi[1]),
         ptr(dgh[0]), ptr(dgh[1]), T, N, st)
    errs = []
    for d, s in enumerate(("", "_reverse")):
        errs.append(rel_l2(dgi[d].sum(0), P64["b_ih" + s].grad))
        errs.append(rel_l2(dgh[d].sum(0), P64["b_hh" + s].grad))
        errs.append(rel_l2(dgi[d].double().t() @ xd.double().reshape(T * N, I), P64["w_ih" + s].grad))
    dx = dgi[0].double() @ D["w_ih"].double() + dgi[1].double() @ D["w_ih_reverse"].double()
    errs.append(rel_l2(dx.reshape(T, N, I), x64.grad))
    print(f"GRU T={T} N={N}: forward rel-L2 {e_f:.2e}; gradients max rel-L2 {max(errs):.2e}")
    assert e_f < 1e-5 and max(errs) < 1e-4


@pytest.mark.parametrize("kind", ["det", "rec"])
def test_eval_mode_backward_uses_running_statistics(kind):
    """model.eval() forward with autograd on (reference test() loops, frozen-BN fine-tuning): BatchNorm backward
    must be dz * gamma * invstd(running), not the batch-statistics formula."""
    from ocrs_models_b200 import CTCLoss, DetectionModel, balanced_cross_entropy_loss

    g = torch.Generator().manual_seed(11)
    torch.manual_seed(1234)
    if kind == "det":
        m = DetectionModel()
        batch = {"image": torch.rand(2, 1, 96, 80, generator=g) - 0.5, "mask": (torch.rand(2, 1, 96, 80, generator=g) < 0.1).float()}
        loss_fn = lambda y: balanced_cross_entropy_loss(y, batch["mask"].cuda())  # noqa: E731
    else:
        m = _rec()
        batch = {"image": torch.rand(3, 1, 64, 96, generator=g) - 0.5, "targets": torch.randint(1, 97, (3, 8), generator=g, dtype=torch.int32),
                 "input_lengths": torch.tensor([24, 24, 20]), "target_lengths": torch.tensor([8, 5, 1])}
        ctc = CTCLoss()
        loss_fn = lambda lp: ctc(lp, batch["targets"].cuda(), batch["input_lengths"], batch["target_lengths"])  # noqa: E731
    with torch.no_grad():  # running statistics away from (0, 1) so the two formulas differ
        for k, b in m.named_buffers():
            if k.endswith("running_mean"):
                b.add_(torch.randn(b.shape, generator=g) * 0.05)
            elif k.endswith("running_var"):
                b.mul_(0.5 + torch.rand(b.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out64, loss64, g64, _ = O.train_step_grads(kind, sd, batch, torch.float64, training=False)
    m = m.cuda().eval()
    out = m(batch["image"].cuda())
    loss = loss_fn(out)
    loss.backward()
    assert rel_l2(out, out64) < 1e-4 and abs(loss.item() - loss64.item()) < 1e-4 * abs(loss64.item())
    gn = torch.sqrt(sum((v ** 2).sum() for v in g64.values()))
    err = float(torch.sqrt(sum(((p.grad.cpu().double() - g64[k]) ** 2).sum() for k, p in m.named_parameters())) / gn)
    print(f"eval-mode backward {kind}: global grad rel-L2 {err:.2e}")
    assert err < 1e-3
    for k, v in m.state_dict().items():  # eval mode must not touch the buffers
        assert torch.equal(v.cpu(), sd[k]), k
