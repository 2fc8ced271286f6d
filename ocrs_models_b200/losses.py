"""Loss functions on the two training hot paths, backed by the sm_100a kernels.

* :class:`CTCLoss` mirrors ``torch.nn.CTCLoss`` as constructed at reference
  ``ocrs_models/train_rec.py:104`` and called at ``:121`` / ``:212``.
* :func:`balanced_cross_entropy_loss` mirrors reference
  ``ocrs_models/train_detection.py:225-263``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from ._lib import call, ptr

_REDUCTION = {"none": 0, "mean": 1, "sum": 2}


def _lengths_to_device(lengths, n: int, device) -> tuple[torch.Tensor, int | None]:
    """int32 device copy of a lengths argument + its host-side max when known without a sync."""
    if isinstance(lengths, torch.Tensor):
        host_max = int(lengths.max()) if (lengths.device.type == "cpu" and lengths.numel()) else None
        t = lengths.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    else:
        lst = [int(v) for v in lengths]
        host_max = max(lst) if lst else 0
        t = torch.tensor(lst, dtype=torch.int32, device=device)
    if t.numel() != n:
        raise RuntimeError(f"lengths must have {n} entries, got {t.numel()}")
    return t, host_max


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, reduction, zero_infinity):
        if not log_probs.is_cuda:
            raise RuntimeError("ocrs_models_b200.CTCLoss has no CPU path: log_probs must be a CUDA tensor")
        if log_probs.dim() != 3:
            raise RuntimeError("log_probs must be (T, N, C)")
        T, N, C = log_probs.shape
        dev = log_probs.device
        lp = log_probs.detach().float().contiguous()
        il, _ = _lengths_to_device(input_lengths, N, dev)
        tl, tl_max = _lengths_to_device(target_lengths, N, dev)
        if targets.dim() == 1:
            # concatenated form -> padded (N, S_max); needs the lengths on the host
            tl_host = target_lengths.tolist() if isinstance(target_lengths, torch.Tensor) else list(target_lengths)
            s_max = max(max(tl_host), 1)
            padded = torch.zeros((N, s_max), dtype=torch.int32)
            off = 0
            tcpu = targets.detach().cpu()
            for i, s in enumerate(tl_host):
                padded[i, :s] = tcpu[off : off + s]
                off += s
            tg = padded.to(dev)
        elif targets.dim() == 2:
            tg = targets.detach().to(device=dev, dtype=torch.int32).contiguous()
        else:
            raise RuntimeError("targets must be 1-D or 2-D")
        max_s = tg.shape[1] if tl_max is None else min(tl_max, tg.shape[1])
        if tl_max is not None and tl_max > tg.shape[1]:
            raise RuntimeError("target_lengths exceed the padded target width")
        row = _lib.lib().ocrs_ctc_alpha_row(max_s)
        if row == 0:
            raise RuntimeError(f"CTC target length {max_s} exceeds the supported maximum of 255")
        alpha = torch.empty((N, T, row), dtype=torch.float32, device=dev)
        nll = torch.empty((N,), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            call("ocrs_ctc_fwd", ptr(lp), ptr(tg), tg.stride(0) if tg.numel() else 0, ptr(il), ptr(tl),
                 T, N, C, max_s, blank, reduction, int(zero_infinity), ptr(alpha), ptr(nll), ptr(loss), st)
        ctx.save_for_backward(lp, tg, il, tl, alpha, nll)
        ctx.cfg = (T, N, C, max_s, blank, reduction, int(zero_infinity))
        if reduction == 0:
            out = nll.clone()
            if zero_infinity:
                out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
            return out
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lp, tg, il, tl, alpha, nll = ctx.saved_tensors
        T, N, C, max_s, blank, reduction, zero_inf = ctx.cfg
        dev = lp.device
        go = grad_out.detach().float().contiguous()
        if reduction == 0 and go.numel() != N:
            go = go.expand(N).contiguous()
        grad = torch.empty_like(lp)
        with torch.cuda.device(dev):
            call("ocrs_ctc_bwd", ptr(lp), ptr(tg), tg.stride(0) if tg.numel() else 0, ptr(il), ptr(tl),
                 T, N, C, max_s, blank, reduction, zero_inf, ptr(alpha), ptr(nll), ptr(go), ptr(grad),
                 _lib.stream_ptr(dev))
        return grad, None, None, None, None, None, None


class CTCLoss(nn.Module):
    """Drop-in for ``torch.nn.CTCLoss`` (same arguments, same gradient convention).

    The gradient returned for ``log_probs`` follows aten's softmax-folded convention
    (``exp(lp) - posterior``), so the gradient that reaches the logits through a preceding
    ``LogSoftmax`` is identical to the reference's.
    """

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _REDUCTION:
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return _CTCFunction.apply(
            log_probs, targets, input_lengths, target_lengths, self.blank, _REDUCTION[self.reduction], self.zero_infinity
        )


class _BalancedBCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        if not pred.is_cuda:
            raise RuntimeError("ocrs_models_b200.balanced_cross_entropy_loss has no CPU path")
        if pred.shape != target.shape:
            raise RuntimeError("pred and target must have the same shape")
        dev = pred.device
        p = pred.detach().float().contiguous()
        t = target.detach().to(device=dev, dtype=torch.float32).contiguous()
        n = p.numel()
        lib = _lib.lib()
        state = torch.zeros((lib.ocrs_bce_state_words(),), dtype=torch.int32, device=dev)
        loss_map = torch.empty((n,), dtype=torch.float32, device=dev)
        partials = torch.empty((lib.ocrs_bce_blocks(),), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("ocrs_balanced_bce_fwd", ptr(p), ptr(t), n, ptr(loss_map), ptr(state), ptr(partials), ptr(loss),
                 _lib.stream_ptr(dev))
        ctx.save_for_backward(p, t, loss_map, state)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        p, t, loss_map, state = ctx.saved_tensors
        dev = p.device
        go = grad_out.detach().float().contiguous()
        dp = torch.empty_like(p)
        with torch.cuda.device(dev):
            call("ocrs_balanced_bce_bwd", ptr(p), ptr(t), ptr(loss_map), p.numel(), ptr(state), ptr(go), ptr(dp),
                 _lib.stream_ptr(dev))
        return dp, None


def balanced_cross_entropy_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Balanced BCE between probability maps (reference train_detection.py:225-263), fully on
    device: no host sync for the positive/negative counts, radix-select instead of two topk."""
    return _BalancedBCEFunction.apply(pred, target)
