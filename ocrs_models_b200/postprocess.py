"""Eval-path post-processing with the heavy part on the device.

:func:`extract_cc_quads` is a drop-in for reference ``ocrs_models/postprocess.py:11-36``: bounding quads of the connected
components of a binary text mask. The reference copies the whole mask to the host and runs ``cv2.findContours`` +
``cv2.minAreaRect`` per component; here the mask is binarised and labelled on the GPU (8-connectivity, ``csrc/
postprocess.cu``), only the components' boundary pixels come back, and ``cv2.minAreaRect`` (a convex hull + rotating
calipers over a few hundred points) gives the same rectangles, because the minimum-area rectangle of a contour depends
only on its convex hull. :func:`connected_components` exposes the label maps; :func:`batch_cc_quads` does a whole batch
of predicted masks with one labelling pass (what ``train_detection.test()``, :144-195, needs per batch).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr


def connected_components(mask: torch.Tensor, threshold: float = 0.5):
    """mask: CUDA tensor [H,W], [1,H,W], [N,H,W] or [N,1,H,W]; foreground = mask > threshold (binarize_mask,
    train_detection.py:33-34). Returns (labels int32 [N,H,W]: 0 = background, else 1 + smallest linear pixel index of the
    8-connected component; n_components int32 [N])."""
    if not mask.is_cuda:
        raise RuntimeError("ocrs_models_b200.postprocess has no CPU path: mask must be a CUDA tensor")
    m = mask.detach().float()
    if m.dim() == 4:
        if m.shape[1] != 1:
            raise ValueError("expected an Nx1xHxW mask")
        m = m[:, 0]
    elif m.dim() == 2:
        m = m[None]
    elif m.dim() != 3:
        raise ValueError("expected an HxW, NxHxW or Nx1xHxW mask")
    m = m.contiguous()
    N, H, W = m.shape
    dev = m.device
    scratch = torch.empty((N, H, W), dtype=torch.int32, device=dev)
    labels = torch.empty((N, H, W), dtype=torch.int32, device=dev)
    ncomp = torch.zeros((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("ocrs_cc_label", ptr(m), float(threshold), N, H, W, ptr(scratch), ptr(labels), ptr(ncomp), _lib.stream_ptr(dev))
    return labels, ncomp


def _boundary_points(labels: torch.Tensor):
    N, H, W = labels.shape
    dev = labels.device
    cap = max(1024, (H * W) // 8)
    while True:
        pts = torch.empty((N, cap, 3), dtype=torch.int32, device=dev)
        cnt = torch.zeros((N,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            call("ocrs_cc_boundary", ptr(labels), N, H, W, ptr(pts), cap, ptr(cnt), _lib.stream_ptr(dev))
        counts = cnt.cpu()
        if int(counts.max()) <= cap:
            return pts, counts
        cap = int(counts.max())


def batch_cc_quads(masks: torch.Tensor, threshold: float = 0.5) -> list[torch.Tensor]:
    """One labelling pass for a batch of masks; returns one Kx4x2 float tensor of quads per image (cv2.boxPoints order)."""
    import cv2

    labels, _ = connected_components(masks, threshold)
    pts, counts = _boundary_points(labels)
    out = []
    for n in range(labels.shape[0]):
        p = pts[n, : int(counts[n])].cpu().numpy()
        quads = []
        if len(p):
            order = np.argsort(p[:, 0], kind="stable")
            p = p[order]
            starts = np.flatnonzero(np.r_[True, p[1:, 0] != p[:-1, 0]])
            for a, b in zip(starts, np.r_[starts[1:], len(p)]):
                quads.append(cv2.boxPoints(cv2.minAreaRect(p[a:b, 1:3].astype(np.int32))))
        out.append(torch.tensor(np.array(quads), dtype=torch.float32).reshape(-1, 4, 2))
    return out


def extract_cc_quads(mask: torch.Tensor) -> torch.Tensor:
    """Drop-in for reference postprocess.extract_cc_quads (HxW or 1xHxW mask; non-zero = text) for CUDA masks."""
    if mask.dim() > 2:
        if mask.shape[0] != 1:
            raise ValueError("Expected mask to be an HxW or 1xHxW tensor")
        mask = mask[0]
    return batch_cc_quads(mask[None].to(torch.uint8).float(), threshold=0.5)[0]
