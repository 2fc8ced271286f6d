"""Recognition accuracy bookkeeping on the device.

:class:`RecognitionAccuracyStats` mirrors the class of the same name in reference
``ocrs_models/train_rec.py:20-82`` (``update`` / ``char_error_rate`` / ``stats_dict``): greedy CTC decoding
(``ocrs_models/datasets/util.py:163-177``) and the Levenshtein distance to the target text run in one kernel
(``csrc/metrics.cu``), the running totals stay in device memory, and nothing synchronises the host until
``char_error_rate()`` is read. ``install()`` rebinds the reference's class to this one, so the unmodified
``train_rec.train()`` / ``test()`` loops stop paying ``.tolist()`` + Python edit distances per batch.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr


def _as_i32(lengths, n: int, device) -> torch.Tensor:
    if isinstance(lengths, torch.Tensor):
        t = lengths.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    else:
        t = torch.tensor([int(v) for v in lengths], dtype=torch.int32, device=device)
    if t.numel() != n:
        raise RuntimeError(f"expected {n} lengths, got {t.numel()}")
    return t


def greedy_decode_cer(preds: torch.Tensor, pred_lengths, targets: torch.Tensor, blank: int = 0,
                      return_decoded: bool = False, total: torch.Tensor | None = None):
    """Per-sample edit distance between the greedy CTC decoding of ``preds`` ([T, N, C] scores or log-probs, CUDA)
    and the non-blank labels of ``targets`` ([N, S_pad] int). Returns int32 ``[N]`` (and, with ``return_decoded``,
    the decoded label sequences ``[N, T]`` + their lengths). ``total``: optional int64 ``[1]`` device accumulator."""
    if not preds.is_cuda:
        raise RuntimeError("ocrs_models_b200.metrics has no CPU path: preds must be a CUDA tensor")
    if preds.dim() != 3:
        raise RuntimeError("preds must be (T, N, C)")
    T, N, C = preds.shape
    dev = preds.device
    lp = preds.detach().float().contiguous()
    if targets.dim() != 2 or targets.shape[0] != N:
        raise RuntimeError("targets must be (N, S_pad)")
    tg = targets.detach().to(device=dev, dtype=torch.int32).contiguous()
    lib = _lib.lib()
    if tg.shape[1] > lib.ocrs_ctc_greedy_cer_max_targets():
        raise RuntimeError(f"target rows longer than {lib.ocrs_ctc_greedy_cer_max_targets()} labels are not supported")
    pl = _as_i32(pred_lengths, N, dev)
    dist = torch.empty((N,), dtype=torch.int32, device=dev)
    dec = torch.zeros((N, T), dtype=torch.int32, device=dev) if return_decoded else None
    dlen = torch.empty((N,), dtype=torch.int32, device=dev) if return_decoded else None
    with torch.cuda.device(dev):
        call("ocrs_ctc_greedy_cer", ptr(lp), T, N, C, ptr(pl), ptr(tg), tg.stride(0) if tg.numel() else 0, tg.shape[1], blank,
             ptr(dist), ptr(dec), ptr(dlen), ptr(total), _lib.stream_ptr(dev))
    return (dist, dec, dlen) if return_decoded else dist


class RecognitionAccuracyStats:
    """Drop-in for reference ``train_rec.RecognitionAccuracyStats`` with device-resident totals."""

    def __init__(self):
        self._errors: torch.Tensor | None = None  # int64 [1] on the device of the first update
        self._chars: torch.Tensor | None = None

    def update(self, targets: torch.Tensor, target_lengths, preds: torch.Tensor, pred_lengths):
        """Same arguments as the reference: targets [batch, seq], target lengths, preds [seq, batch, class], pred lengths."""
        assert len(target_lengths) == targets.size(0)
        assert len(pred_lengths) == preds.size(1)
        dev = preds.device
        if self._errors is None:
            self._errors = torch.zeros(1, dtype=torch.int64, device=dev)
            self._chars = torch.zeros(1, dtype=torch.int64, device=dev)
        greedy_decode_cer(preds, pred_lengths, targets, total=self._errors)
        tl = target_lengths if isinstance(target_lengths, torch.Tensor) else torch.tensor([int(v) for v in target_lengths])
        self._chars += tl.to(device=dev, dtype=torch.int64, non_blocking=True).sum()

    @property
    def total_chars(self) -> int:
        return 0 if self._chars is None else int(self._chars.item())

    @property
    def char_errors(self) -> int:
        return 0 if self._errors is None else int(self._errors.item())

    def char_error_rate(self) -> float:
        """Overall fraction of character-level errors (the one place that synchronises with the device)."""
        return self.char_errors / self.total_chars

    def stats_dict(self) -> dict:
        return {"char_error_rate": self.char_error_rate()}
