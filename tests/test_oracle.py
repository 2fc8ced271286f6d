"""The oracle (oracle/functional.py) against golden vectors recorded from the reference itself."""
import numpy as np
import pytest
import torch

from ocrs_models_b200 import DetectionModel, RecognitionModel
from oracle import functional as O


def _stat(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()])


def _check_stats(gold, prefix, named, rtol, atol_scale=0.0):
    keys = [k for k in gold if k.startswith(prefix + ".stat.")]
    assert keys
    for k in keys:
        name = k[len(prefix) + 6 :]
        got, want = _stat(named[name]), gold[k]
        atol = atol_scale * want[2]
        assert np.allclose(got[1:], want[1:], rtol=rtol, atol=atol), (name, got, want)
        assert abs(got[0] - want[0]) <= rtol * want[1] + atol, (name, got, want)
        full = prefix + ".full." + name
        if full in gold:
            np.testing.assert_allclose(named[name].detach().float().numpy(), gold[full], rtol=rtol * 10, atol=max(atol, rtol * want[2]))


def _det_batch(n, h, w):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 1, h, w, generator=g) - 0.5
    m = (torch.rand(n, 1, h, w, generator=g) < 0.1).float()
    return {"image": x, "mask": m}


def _rec_batch(gold):
    n, w, s_pad = gold["meta"]
    g = torch.Generator().manual_seed(0)
    x = torch.rand(int(n), 1, 64, int(w), generator=g) - 0.5
    return {
        "image": x,
        "targets": torch.from_numpy(gold["targets"]),
        "input_lengths": torch.from_numpy(gold["il"]),
        "target_lengths": torch.from_numpy(gold["tl"]),
    }


def test_param_tree_matches_reference_init(golden):
    """Same names, shapes and seed-1234 initial values as the reference constructors."""
    gold = golden("det_96x80")
    torch.manual_seed(1234)
    det = DetectionModel()
    _check_stats(gold, "param", dict(det.named_parameters()), rtol=1e-6)
    assert sum(p.numel() for p in det.parameters()) == 622122
    gold = golden("rec_w96")
    torch.manual_seed(1234)
    rec = RecognitionModel(O.DEFAULT_ALPHABET)
    _check_stats(gold, "param", dict(rec.named_parameters()), rtol=1e-6)
    assert sum(p.numel() for p in rec.parameters()) == 2426913
    assert len(O.DEFAULT_ALPHABET) == 96


@pytest.mark.parametrize("case,dtype", [("det_96x80", torch.float32), ("det_96x80_fp64", torch.float64), ("det_kat_256", torch.float32)])
def test_det_oracle_vs_reference(golden, case, dtype):
    gold = golden(case)
    n, h, w = (int(v) for v in gold["meta"])
    torch.manual_seed(1234)
    sd = DetectionModel().state_dict()
    out, loss, grads, nb = O.train_step_grads("det", sd, _det_batch(n, h, w), dtype)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    assert abs(loss.item() - gold["loss"]) <= max(tol, 1e-6) * abs(gold["loss"])
    assert np.allclose([out.mean().item(), out.std().item()], gold["y_stats"], rtol=1e-5)
    if "y" in gold:
        np.testing.assert_allclose(out.float().numpy(), gold["y"], rtol=1e-4, atol=1e-6)
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    assert abs(gn - gold["grad_norm"]) <= (1e-8 if dtype == torch.float64 else 1e-3) * gold["grad_norm"]
    if dtype == torch.float64:
        _check_stats(gold, "grad", grads, rtol=1e-7, atol_scale=1e-9)
        _check_stats(gold, "buf", {k: v for k, v in nb.items() if v.is_floating_point()}, rtol=1e-6)
    else:
        # fp32 vs fp32 with a different op order: per-tensor agreement up to the conditioning noise floor
        _check_stats(gold, "buf", {k: v for k, v in nb.items() if v.is_floating_point()}, rtol=1e-4, atol_scale=1e-5)


def test_det_kat_values_from_survey(golden):
    """SURVEY.md section 8c KAT-DET known answers (recorded independently of make_golden.py)."""
    gold = golden("det_kat_256")
    assert abs(gold["loss"] - 0.81453931) < 2e-6
    assert abs(gold["grad_norm"] - 0.314683) < 2e-5
    assert int(gold["n_pos"]) == 13167
    assert np.allclose(gold["y_stats"], [0.44807869, 0.07055192], atol=2e-6)


@pytest.mark.parametrize("case,dtype", [("rec_w96", torch.float32), ("rec_w96_fp64", torch.float64), ("rec_kat_800", torch.float32)])
def test_rec_oracle_vs_reference(golden, case, dtype):
    gold = golden(case)
    torch.manual_seed(1234)
    sd = RecognitionModel(O.DEFAULT_ALPHABET).state_dict()
    out, loss, grads, nb = O.train_step_grads("rec", sd, _rec_batch(gold), dtype)
    tol = 1e-9 if dtype == torch.float64 else 1e-4
    assert abs(loss.item() - gold["loss"]) <= tol * abs(gold["loss"])
    assert np.allclose([out.mean().item(), out.std().item()], gold["lp_stats"], rtol=1e-5)
    if "lp" in gold:
        np.testing.assert_allclose(out.float().numpy(), gold["lp"], rtol=1e-4, atol=1e-5)
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    assert abs(gn - gold["grad_norm"]) <= (1e-8 if dtype == torch.float64 else 1e-3) * gold["grad_norm"]
    if dtype == torch.float64:
        _check_stats(gold, "grad", grads, rtol=1e-7, atol_scale=1e-9)
        _check_stats(gold, "buf", {k: v for k, v in nb.items() if v.is_floating_point()}, rtol=1e-6)


def test_ctc_numpy_restatement_vs_aten(golden):
    gold = golden("ctc_small")
    lp, tgt, il, tl = gold["lp"], gold["targets"], gold["il"], gold["tl"]
    for n in range(lp.shape[1]):
        nll, grad = O.ctc_nll_numpy(lp[:, n], tgt[n, : tl[n]], int(il[n]))
        assert abs(nll - gold["loss_none"][n]) < 1e-5
        np.testing.assert_allclose(grad, gold["grad_none"][:, n], atol=1e-5)
    per = gold["loss_none"] / np.maximum(tl, 1)
    assert abs(per.mean() - gold["loss_mean"]) < 1e-6


def test_adam_and_clip_restatement():
    torch.manual_seed(0)
    p = {"a": torch.randn(7, 5), "b": torch.randn(11)}
    ref = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.Adam(ref.values())
    state: dict = {}
    for _ in range(3):
        g = {k: torch.randn_like(v) * 3 for k, v in p.items()}
        for k in ref:
            ref[k].grad = g[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_(ref.values(), 4.0)
        opt.step()
        n = O.clip_grad_norm(g, 4.0)
        O.adam_step(p, g, state)
        assert abs(n - n_ref) < 1e-5
        for k in p:
            torch.testing.assert_close(p[k], ref[k].detach(), rtol=1e-5, atol=1e-6)


def test_bench_first_step_constants_present():
    """bench.py asserts its first-step losses against these reference-derived constants (oracle/bench_constants.py)."""
    import json
    import os

    from conftest import GOLDEN

    c = json.load(open(os.path.join(GOLDEN, "bench_first_step.json")))
    assert np.isfinite([c["rec_loss_f64"], c["det_loss_f64"]]).all()
    assert abs(c["rec_loss_f32"] - c["rec_loss_f64"]) < 1e-4 * c["rec_loss_f64"]
    assert abs(c["det_loss_f32"] - c["det_loss_f64"]) < 1e-4 * c["det_loss_f64"]


def test_greedy_cer_oracle_matches_the_reference_functions():
    """oracle.greedy_cer vs the reference's own decode_text / ctc_greedy_decode_text (datasets/util.py) + Levenshtein."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from baseline import ref_loader

    if not ref_loader.available():
        pytest.skip("baseline/_ref not staged")
    ref_loader.load()
    from ocrs_models.datasets.util import ctc_greedy_decode_text, decode_text
    from ocrs_models.train_rec import levenshtein  # pylev (or its stand-in under baseline/stubs)

    alphabet = list(O.DEFAULT_ALPHABET)
    g = torch.Generator().manual_seed(0)
    T, N, C, S = 50, 6, 97, 16
    lp = torch.randn(T, N, C, generator=g)
    runs = torch.randint(0, C, (T // 2, N), generator=g).repeat_interleave(2, dim=0)
    lp.scatter_(2, runs.unsqueeze(-1), 8.0)
    tg = torch.randint(0, C, (N, S), generator=g, dtype=torch.int32)
    pl = torch.tensor([50, 40, 0, 13, 50, 1])
    dists, decs = O.greedy_cer(lp, pl, tg)
    labels = lp.argmax(-1).transpose(0, 1).tolist()
    for n in range(N):
        t_text, p_text = decode_text(tg[n].tolist(), alphabet), ctc_greedy_decode_text(labels[n][: int(pl[n])], alphabet)
        assert "".join(alphabet[c - 1] for c in decs[n]) == p_text
        assert dists[n] == levenshtein(t_text, p_text)
    assert O.levenshtein("kitten", "sitting") == 3 and O.levenshtein("", "abc") == 3
