"""Delivery of parameter gradients from the backward kernels: one multi-tensor launch (csrc/glue.cu).

The backward kernels leave every parameter gradient as per-CTA / split-K partial rows. ``deliver`` reduces all of them in
ONE launch of ``ocrs_grad_deliver`` (double-precision sums in a fixed order, like ``ocrs_finalize_partials``) and

* for parameters whose ``.grad`` is the view of a flat gradient bucket installed by ``optim.FusedAdam`` (the parameter
  carries ``_ocrs_grad_sink``), ACCUMULATES straight into that view and hands ``None`` to autograd - the same
  ``grad += g`` autograd's AccumulateGrad would do (reference: ``loss.backward()`` at ocrs_models/train_rec.py:130,
  ocrs_models/train_detection.py:96), without one elementwise kernel per parameter;
* for every other parameter, stores into a fresh tensor that is returned to autograd as usual.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class Partial:
    """A gradient still in partial form: sum over ``rows`` of ``src[r * ld + col(k)]``, k < K.

    ``conv=(Cin, kh*kw)``: the rows are a convolution gradient laid out [Cout][(ky,kx,ci)] (the GEMM's layout) and the
    destination is the parameter layout [Cout][Cin][kh][kw]. ``inner=(run, pitch)``: runs of ``run`` elements at ``pitch``.
    ``off``: element offset of the first column inside a row."""

    __slots__ = ("t", "rows", "K", "ld", "off", "map", "p0", "p1")

    def __init__(self, t: torch.Tensor, rows: int, K: int, ld: int | None = None, off: int = 0, conv=None, inner=None):
        self.t, self.rows, self.K, self.ld, self.off = t, int(rows), int(K), int(K if ld is None else ld), int(off)
        self.map, self.p0, self.p1 = 0, 1, 1
        if conv is not None:
            self.map, self.p0, self.p1 = 1, int(conv[0]), int(conv[1])
        elif inner is not None:
            self.map, self.p0, self.p1 = 2, int(inner[0]), int(inner[1])


def deliver(params, grads: dict, st) -> list:
    """``grads``: id(param) -> Partial | Tensor (finished, contiguous) | None. Returns the list autograd expects, in the
    order of ``params`` (None where the gradient went straight into the parameter's gradient bucket)."""
    lib = _lib.lib()
    out, entries, keep = [], [], []
    for p in params:
        g = grads.get(id(p))
        if g is None:
            out.append(None)
            continue
        if isinstance(g, torch.Tensor):
            g = Partial(g, 1, g.numel())
        if g.K != p.numel():
            raise RuntimeError(f"gradient of {tuple(p.shape)} has {g.K} elements")
        sink = getattr(p, "_ocrs_grad_sink", None)
        if sink is not None and p.grad is not None and p.grad.data_ptr() == sink.data_ptr() and p.grad.is_contiguous():
            dst, acc = sink, 1
            out.append(None)
        else:
            dst, acc = torch.empty_like(p, memory_format=torch.contiguous_format), 0
            out.append(dst)
        entries.append((g.t.data_ptr() + 4 * g.off, dst.data_ptr(), g.K, g.rows, g.ld, g.map, g.p0, g.p1, acc))
        keep.append(g.t)
    step = lib.ocrs_grad_deliver_max()
    for i in range(0, len(entries), step):
        chunk = entries[i : i + step]
        n = len(chunk)
        cols = list(zip(*chunk))
        src = (ctypes.c_void_p * n)(*cols[0])
        dst = (ctypes.c_void_p * n)(*cols[1])
        ints = [(ctypes.c_int * n)(*c) for c in cols[2:]]
        _lib.call("ocrs_grad_deliver", src, dst, *ints, n, st)
    return out


def materialize(g, shape, st) -> torch.Tensor:
    """The finished gradient of one Partial as a new tensor of `shape` (tests, and callers outside a backward pass)."""
    if isinstance(g, torch.Tensor):
        return g.reshape(shape)
    out = torch.empty(shape, dtype=torch.float32, device=g.t.device)
    assert out.numel() == g.K
    one = lambda T, v: (T * 1)(v)
    _lib.call("ocrs_grad_deliver", one(ctypes.c_void_p, g.t.data_ptr() + 4 * g.off), one(ctypes.c_void_p, out.data_ptr()),
              one(ctypes.c_int, g.K), one(ctypes.c_int, g.rows), one(ctypes.c_int, g.ld), one(ctypes.c_int, g.map),
              one(ctypes.c_int, g.p0), one(ctypes.c_int, g.p1), one(ctypes.c_int, 0), 1, st)
    return out
