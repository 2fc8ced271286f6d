"""CPU oracle: a functional restatement of the reference's two training hot paths.

TEST INFRASTRUCTURE ONLY. Nothing under ``ocrs_models_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs do, and only as the checker or the timed CPU baseline.

The reference (robertknight/ocrs-models @ 3d98fc6) is pure Python on top of PyTorch; the
arithmetic lives in the third-party, *unpinned* ``torch`` dependency (absent from
pyproject.toml / poetry.lock). This restatement therefore expresses the same algorithm with
``torch.nn.functional`` CPU ops (version in this image: torch 2.11.0) driven by a plain
``state_dict`` with the reference's parameter names, in any float dtype (fp64 = ground truth).
It is pinned against the reference itself: ``oracle/make_golden.py`` imports
``/root/reference/ocrs_models`` in the build container and records outputs, losses and
gradients under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against them.

Each function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DET_DEPTH_SCALE = [8, 16, 32, 32, 64, 128, 256]  # ocrs_models/models.py:112
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# ocrs_models/datasets/hiertext.py:133-137
DEFAULT_ALPHABET = (
    " 0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~" + chr(8364) + "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
)


def _bn(sd, prefix, x, training, new_buffers):
    """nn.BatchNorm2d (models.py:23,197,214,231,241): eps 1e-5, momentum 0.1, unbiased running var."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        dims = (0, 2, 3)
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        n = x.numel() // x.shape[1]
        if new_buffers is not None:
            with torch.no_grad():
                new_buffers[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach().to(rm.dtype)
                unb = var.detach() * (n / max(n - 1, 1))
                new_buffers[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unb.to(rv.dtype)
                new_buffers[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    xhat = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    return xhat * w[None, :, None, None] + b[None, :, None, None]


def _depthwise_block(sd, prefix, x, training, nb):
    """DepthwiseConv (models.py:7-28): dw3x3(pad 1, no bias) -> 1x1(no bias) -> BN -> ReLU."""
    c = x.shape[1]
    x = F.conv2d(x, sd[prefix + ".seq.0.weight"], None, padding=1, groups=c)
    x = F.conv2d(x, sd[prefix + ".seq.1.weight"], None)
    x = _bn(sd, prefix + ".seq.2", x, training, nb)
    return F.relu(x)


def _double_conv(sd, prefix, x, training, nb):
    """DoubleConv (models.py:31-41)."""
    x = _depthwise_block(sd, prefix + ".seq.0", x, training, nb)
    return _depthwise_block(sd, prefix + ".seq.1", x, training, nb)


def det_forward(sd: dict, x: torch.Tensor, training: bool = True, new_buffers: dict | None = None) -> torch.Tensor:
    """DetectionModel.forward (models.py:131-143) -> per-pixel text probability."""
    n_levels = len(DET_DEPTH_SCALE) - 1
    x = _double_conv(sd, "in_conv", x, training, new_buffers)
    downs = []
    for i in range(n_levels):  # Down (models.py:44-58)
        prev = x if i == 0 else downs[-1]
        y = _double_conv(sd, f"down.{i}.seq.0", prev, training, new_buffers)
        downs.append(F.max_pool2d(y, 2))
    up = downs[-1]
    for i in reversed(range(n_levels)):  # Up (models.py:61-90)
        skip = x if i == 0 else downs[i - 1]
        u = F.conv_transpose2d(up, sd[f"up.{i}.up.weight"], sd[f"up.{i}.up.bias"], stride=2)
        u = u[:, :, : skip.shape[2], : skip.shape[3]]
        up = _double_conv(sd, f"up.{i}.contract", torch.cat((u, skip), dim=1), training, new_buffers)
    z = F.conv2d(up, sd["out_conv.0.weight"], sd["out_conv.0.bias"])  # models.py:126-129
    return torch.sigmoid(z)


def balanced_cross_entropy_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """train_detection.py:225-263: mean of the top-k masked BCE values of each class, k = min(#pos, #neg)."""
    pos = target > 0.5
    neg = target < 0.5
    t = target.clamp(0.0, 1.0)
    # F.binary_cross_entropy clamps each log term at -100 (torch/nn/modules/loss.py)
    logp = torch.clamp(torch.log(pred), min=-100.0)
    log1mp = torch.clamp(torch.log(1 - pred), min=-100.0)
    pixel = -(t * logp + (1 - t) * log1mp)
    k = int(min(int(pos.sum()), int(neg.sum())))
    pos_top = torch.topk((pos * pixel).flatten(), k, sorted=False).values
    neg_top = torch.topk((neg * pixel).flatten(), k, sorted=False).values
    return torch.cat([pos_top, neg_top]).mean()


# ------------------------------------------------------------------------------------------
# recognition


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One direction of nn.GRU (models.py:245), h0 = 0. Gate order (r, z, n):
    r = s(gi_r + gh_r); z = s(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1 - z) * n + z * h."""
    T, N, _ = x.shape
    H = w_hh.shape[1]
    gi_all = x @ w_ih.t() + b_ih
    h = x.new_zeros((N, H))
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
        nn_ = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
        h = (1 - z) * nn_ + z * h
        outs[t] = h
    return torch.stack(outs, 0)


def gru_forward(sd: dict, x: torch.Tensor, prefix: str = "gru", num_layers: int = 2) -> torch.Tensor:
    """2-layer bidirectional GRU (models.py:245,264-266); layer l+1 consumes concat(fwd, rev)."""
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            outs.append(
                _gru_direction(
                    x,
                    sd[f"{prefix}.weight_ih_l{layer}{suffix}"],
                    sd[f"{prefix}.weight_hh_l{layer}{suffix}"],
                    sd[f"{prefix}.bias_ih_l{layer}{suffix}"],
                    sd[f"{prefix}.bias_hh_l{layer}{suffix}"],
                    rev,
                )
            )
        x = torch.cat(outs, dim=2)
    return x


def rec_conv_stack(sd: dict, x: torch.Tensor, training: bool = True, new_buffers: dict | None = None) -> torch.Tensor:
    """RecognitionModel.conv (models.py:179-243)."""
    x = F.max_pool2d(F.relu(F.conv2d(x, sd["conv.0.weight"], sd["conv.0.bias"], padding=1)), 2)
    x = F.conv2d(x, sd["conv.3.weight"], None, padding=1)
    x = F.max_pool2d(F.relu(_bn(sd, "conv.4", x, training, new_buffers)), 2)
    x = F.relu(F.conv2d(x, sd["conv.7.weight"], sd["conv.7.bias"], padding=1))
    x = F.conv2d(x, sd["conv.9.weight"], None, padding=1)
    x = F.max_pool2d(F.relu(_bn(sd, "conv.10", x, training, new_buffers)), (2, 1))
    x = F.relu(F.conv2d(x, sd["conv.13.weight"], sd["conv.13.bias"], padding=1))
    x = F.conv2d(x, sd["conv.15.weight"], None, padding=1)
    x = F.max_pool2d(F.relu(_bn(sd, "conv.16", x, training, new_buffers)), (2, 1))
    x = F.conv2d(x, sd["conv.19.weight"], None, padding=1)
    x = _bn(sd, "conv.20", x, training, new_buffers)
    return F.avg_pool2d(x, (4, 1))


def rec_forward(sd: dict, x: torch.Tensor, training: bool = True, new_buffers: dict | None = None) -> torch.Tensor:
    """RecognitionModel.forward (models.py:253-268) -> (W//4 + 1, N, classes) log-probs."""
    x = rec_conv_stack(sd, x, training, new_buffers)
    x = torch.permute(x, (3, 0, 1, 2))
    x = torch.reshape(x, (x.shape[0], x.shape[1], -1))
    x = gru_forward(sd, x)
    x = x @ sd["output.0.weight"].t() + sd["output.0.bias"]
    return F.log_softmax(x, dim=2)


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False):
    """torch.nn.CTCLoss() as used at train_rec.py:104,121 (aten ctc_loss on CPU)."""
    return F.ctc_loss(log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity)


def ctc_nll_numpy(log_probs: np.ndarray, target: np.ndarray, input_len: int, blank: int = 0):
    """Independent fp64 restatement of the CTC lattice for ONE sample (small cases only).

    Returns (nll, grad) where grad follows aten's convention d/dlp = exp(lp) - posterior
    (Graves et al. 2006, eq. 16, with the softmax folded in), zero for t >= input_len.
    """
    lp = np.asarray(log_probs, dtype=np.float64)
    T, C = lp.shape
    S = len(target)
    L = 2 * S + 1
    lab = [blank if s % 2 == 0 else int(target[s // 2]) for s in range(L)]
    ninf = -math.inf

    def lse(*v):
        m = max(v)
        if m == ninf:
            return ninf
        return m + math.log(sum(math.exp(a - m) for a in v))

    Tn = input_len
    alpha = np.full((Tn, L), ninf)
    beta = np.full((Tn, L), ninf)
    alpha[0, 0] = lp[0, blank]
    if L > 1:
        alpha[0, 1] = lp[0, lab[1]]
    for t in range(1, Tn):
        for s in range(L):
            a = [alpha[t - 1, s]]
            if s >= 1:
                a.append(alpha[t - 1, s - 1])
            if s >= 2 and lab[s] != blank and lab[s] != lab[s - 2]:
                a.append(alpha[t - 1, s - 2])
            alpha[t, s] = lse(*a) + lp[t, lab[s]]
    nll = -lse(alpha[Tn - 1, L - 1], alpha[Tn - 1, L - 2] if L > 1 else ninf)
    beta[Tn - 1, L - 1] = lp[Tn - 1, blank]
    if L > 1:
        beta[Tn - 1, L - 2] = lp[Tn - 1, lab[L - 2]]
    for t in range(Tn - 2, -1, -1):
        for s in range(L):
            b = [beta[t + 1, s]]
            if s + 1 < L:
                b.append(beta[t + 1, s + 1])
            if s + 2 < L and lab[s] != blank and lab[s] != lab[s + 2]:
                b.append(beta[t + 1, s + 2])
            beta[t, s] = lse(*b) + lp[t, lab[s]]
    grad = np.zeros((T, C))
    for t in range(Tn):
        post = np.zeros(C)
        for s in range(L):
            v = alpha[t, s] + beta[t, s]
            if v > ninf:
                post[lab[s]] += math.exp(v + nll - lp[t, lab[s]])
        grad[t] = np.exp(lp[t]) - post
    return nll, grad


# ------------------------------------------------------------------------------------------
# recognition accuracy bookkeeping (train_rec.py:29-68, datasets/util.py:132-177, pylev.levenshtein)


def ctc_greedy_decode_labels(frame_labels) -> list[int]:
    """datasets/util.py:163-177 on label level: skip repeats of the previous frame's label, then blanks (0)."""
    out, last = [], None
    for c in frame_labels:
        if c == last:
            continue
        last = c
        if c == 0:
            continue
        out.append(int(c))
    return out


def levenshtein(a, b) -> int:
    """Edit distance (insert / delete / substitute, unit costs) as computed by pylev.levenshtein (train_rec.py:64)."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def greedy_cer(preds: torch.Tensor, pred_lengths, targets: torch.Tensor):
    """RecognitionAccuracyStats.update (train_rec.py:29-68) per sample: ([edit distances], [decoded label lists]).
    preds [T, N, C]; targets [N, S_pad] (decode_text, util.py:135-147, keeps every label > 0 of the padded row)."""
    labels = preds.argmax(-1).transpose(0, 1).tolist()
    dists, decs = [], []
    for y, x, x_len in zip(targets.tolist(), labels, [int(v) for v in pred_lengths]):
        tgt = [c for c in y if c > 0]
        dec = ctc_greedy_decode_labels(x[:x_len])
        dists.append(levenshtein(tgt, dec))
        decs.append(dec)
    return dists, decs


# ------------------------------------------------------------------------------------------
# optimiser glue (train_detection.py:378, train_rec.py:148,381-382)


def adam_step(params: dict, grads: dict, state: dict, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults (no weight decay, no amsgrad), in place on `params`."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        m = state.setdefault("m." + k, torch.zeros_like(p))
        v = state.setdefault("v." + k, torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2**t)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1**t))


def clip_grad_norm(grads: dict, max_norm: float) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ (train_rec.py:148): scale by max_norm / (norm + 1e-6), capped at 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return total


def param_names(sd: dict) -> list[str]:
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def train_step_grads(kind: str, sd: dict, batch: dict, dtype=torch.float32, training: bool = True):
    """fwd + loss + bwd of one training step on CPU. Returns (output, loss, grads, new_buffers).
    `training=False`: model.eval() forward (running BatchNorm statistics) with autograd still on, as the
    reference's test() loops and frozen-BN fine-tuning do."""
    p = {}
    for k, v in sd.items():
        if v.is_floating_point():
            v = v.detach().to(dtype).clone()
            if k in param_names(sd):
                v.requires_grad_(True)
        p[k] = v
    nb: dict = {}
    if kind == "det":
        out = det_forward(p, batch["image"].to(dtype), training, nb)
        loss = balanced_cross_entropy_loss(out, batch["mask"].to(dtype))
    else:
        out = rec_forward(p, batch["image"].to(dtype), training, nb)
        loss = ctc_loss(out, batch["targets"], batch["input_lengths"], batch["target_lengths"])
    names = param_names(sd)
    gs = torch.autograd.grad(loss, [p[k] for k in names])
    return out.detach(), loss.detach(), dict(zip(names, gs)), nb
