"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run here (build container; /root/reference is absent on the GPU box):
    python oracle/make_golden.py

Imports ``ocrs_models.models`` / ``train_detection`` / ``train_rec`` from /root/reference
(``shapely`` and ``pylev`` are off-path imports that are not installed: stubbed in
sys.modules), builds the models under ``torch.manual_seed(1234)`` as the scripts do
(train_detection.py:337-338), and records outputs, losses, gradients and post-step BN buffers.
Large gradient tensors are stored as (sum, abs-sum, L2) statistics, small ones in full.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("OCRS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FULL_LIMIT = 4096


def import_reference():
    for name, attrs in {
        "shapely": [],
        "shapely.geometry": ["MultiLineString", "JOIN_STYLE", "Polygon"],
        "shapely.geometry.polygon": ["LinearRing", "Polygon"],
        "pylev": ["levenshtein"],
    }.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    sys.path.insert(0, REF)
    from ocrs_models import models, train_detection, train_rec  # noqa
    from ocrs_models.datasets.hiertext import DEFAULT_ALPHABET

    return models, train_detection, train_rec, DEFAULT_ALPHABET


def stats(prefix: str, named: dict, out: dict):
    for k, v in named.items():
        v = v.detach().double()
        out[f"{prefix}.stat.{k}"] = np.array([v.sum().item(), v.abs().sum().item(), v.norm().item()])
        if v.numel() <= FULL_LIMIT:
            out[f"{prefix}.full.{k}"] = v.float().numpy()


def det_case(models, train_detection, name, n, h, w, dtype=torch.float32):
    torch.manual_seed(1234)
    model = models.DetectionModel().train().to(dtype)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 1, h, w, generator=g) - 0.5
    m = (torch.rand(n, 1, h, w, generator=g) < 0.1).float()
    out: dict = {"meta": np.array([n, h, w])}
    stats("param", dict(model.named_parameters()), out)
    y = model(x.to(dtype))
    loss = train_detection.balanced_cross_entropy_loss(y, m.to(dtype))
    loss.backward()
    out["y_stats"] = np.array([y.mean().item(), y.std().item()])
    if y.numel() <= 64 * 1024:
        out["y"] = y.detach().float().numpy()
    out["loss"] = np.array(loss.item())
    out["n_pos"] = np.array(int((m > 0.5).sum()))
    grads = {k: p.grad for k, p in model.named_parameters()}
    out["grad_norm"] = np.array(torch.sqrt(sum((gr.double() ** 2).sum() for gr in grads.values())).item())
    stats("grad", grads, out)
    stats("buf", {k: b for k, b in model.named_buffers() if b.is_floating_point()}, out)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "grad_norm", out["grad_norm"])


def rec_case(models, alphabet, name, n, w, s_pad, tl, il, dtype=torch.float32):
    torch.manual_seed(1234)
    model = models.RecognitionModel(alphabet).train()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 1, 64, w, generator=g) - 0.5
    tgt = torch.randint(1, 97, (n, s_pad), generator=g, dtype=torch.int32)
    tgt[0, 1] = tgt[0, 0]  # a repeated label forces a blank between them
    il_t = torch.tensor(il, dtype=torch.int64)
    tl_t = torch.tensor(tl, dtype=torch.int64)
    out: dict = {"meta": np.array([n, w, s_pad]), "targets": tgt.numpy(), "il": il_t.numpy(), "tl": tl_t.numpy()}
    stats("param", dict(model.named_parameters()), out)
    if dtype == torch.float64:
        # models.py:266 hard-codes x.float(); run the three stages by hand for the fp64 ground truth
        model = model.double()
        f = model.conv(x.double())
        f = torch.permute(f, (3, 0, 1, 2)).reshape(f.shape[3], f.shape[0], -1)
        f, _ = model.gru(f)
        lp = model.output(f)
    else:
        lp = model(x)
    loss = torch.nn.CTCLoss()(lp, tgt, il_t, tl_t)
    loss.backward()
    out["lp_stats"] = np.array([lp.mean().item(), lp.std().item()])
    if lp.numel() <= 64 * 1024:
        out["lp"] = lp.detach().float().numpy()
    out["loss"] = np.array(loss.item())
    grads = {k: p.grad for k, p in model.named_parameters()}
    out["grad_norm"] = np.array(torch.sqrt(sum((gr.double() ** 2).sum() for gr in grads.values())).item())
    stats("grad", grads, out)
    stats("buf", {k: b for k, b in model.named_buffers() if b.is_floating_point()}, out)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "grad_norm", out["grad_norm"])


def ctc_case(name):
    g = torch.Generator().manual_seed(7)
    T, N, C, S = 12, 4, 6, 5
    lp = torch.log_softmax(torch.randn(T, N, C, generator=g), dim=2).requires_grad_(True)
    tgt = torch.tensor([[1, 1, 2, 3, 3], [4, 5, 1, 0, 0], [2, 0, 0, 0, 0], [0, 0, 0, 0, 0]], dtype=torch.int32)
    il = torch.tensor([12, 10, 7, 3])
    tl = torch.tensor([5, 3, 1, 0])
    out = {"lp": lp.detach().numpy(), "targets": tgt.numpy(), "il": il.numpy(), "tl": tl.numpy()}
    for red in ("mean", "sum", "none"):
        lp.grad = None
        loss = torch.nn.CTCLoss(reduction=red)(lp, tgt, il, tl)
        loss.sum().backward()
        out["loss_" + red] = loss.detach().numpy()
        out["grad_" + red] = lp.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, out["loss_mean"])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    models, train_detection, train_rec, alphabet = import_reference()
    ctc_case("ctc_small")
    det_case(models, train_detection, "det_96x80", 2, 96, 80)
    det_case(models, train_detection, "det_96x80_fp64", 2, 96, 80, torch.float64)
    det_case(models, train_detection, "det_kat_256", 2, 256, 256)  # SURVEY 8c KAT-DET
    rec_case(models, alphabet, "rec_w96", 3, 96, 8, [8, 5, 1], [24, 24, 20])
    rec_case(models, alphabet, "rec_w96_fp64", 3, 96, 8, [8, 5, 1], [24, 24, 20], torch.float64)
    rec_case(models, alphabet, "rec_kat_800", 4, 800, 40, [40, 33, 17, 5], [200] * 4)  # SURVEY 8c KAT-REC


if __name__ == "__main__":
    main()
