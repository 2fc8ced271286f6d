"""Execution plan of the detection U-Net on the sm_100a kernels (csrc/det_tma.cu: TMA / cp.async-pipelined blocks of the
8-32 channel levels; csrc/gemm_tc.cu: batched tcgen05 GEMMs of the >= 64 channel levels; csrc/det_convt.cu: ConvTranspose2d
tile kernels; csrc/det_fwd.cu, det_bwd.cu: pooling, out_conv, BatchNorm finalisation and the kernels for shapes TMA cannot
address).

Follows reference ``DetectionModel.forward`` (ocrs_models/models.py:131-143) op for op, but:

* activations are planar NCHW fp32 *views* (tensor, element offset, per-sample stride), so the
  ``torch.cat`` of ``Up.forward`` (models.py:89) is two producers writing into one buffer;
* BatchNorm+ReLU of a block is folded into the loads of its consumers (per-channel scale/shift/lo);
* the whole forward is one ``torch.autograd.Function`` whose backward runs the hand-written
  backward kernels; parameter gradients leave them as partial rows and are reduced and delivered (returned to autograd,
  or accumulated into an attached ``optim.FusedAdam`` bucket) by one multi-tensor launch (``grads.deliver``).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr
from .grads import Partial, deliver

NEG_INF = float("-inf")
# TMA-pipelined DepthwiseConv kernels (csrc/det_tma.cu) wherever the views are TMA-addressable; OCRS_DET_TMA=0 keeps the
# synchronous tile kernels of det_fwd.cu / det_bwd.cu everywhere (A/B testing; they remain the path for W % 4 != 0).
USE_TMA = os.environ.get("OCRS_DET_TMA", "1") == "1"
# BatchNorm-backward sums produced by the kernel that writes d_a (max-pool backward, out_conv backward, depthwise backward)
# instead of a separate pass over d_a and y; OCRS_DET_FUSE_BN=0 restores the separate ocrs_bnrelu_bwd_reduce everywhere.
FUSE_BN_REDUCE = os.environ.get("OCRS_DET_FUSE_BN", "1") == "1"
# 1x1 data gradient computed inside the weight-gradient kernel for blocks with <= 16 output channels (OCRS_DET_FUSE_PWT=0: separate pass).
FUSE_PWT = os.environ.get("OCRS_DET_FUSE_PWT", "1") == "1"
# Levels with >= 64 channels: depthwise 3x3 as its own kernel, the 1x1 convolution and both its gradients as batched
# tcgen05 GEMMs (csrc/gemm_tc.cu); ConvTranspose2d with >= 64 input channels as GEMM + col2im / im2col + GEMMs.
# OCRS_DET_TC=0 keeps the CUDA-core kernels there too.
USE_TC_DEEP = os.environ.get("OCRS_DET_TC", "1") == "1"
# Keep each block's depthwise output from the forward pass for its 1x1 weight gradient (OCRS_DET_SAVE_DW=0: recompute it).
SAVE_DW = os.environ.get("OCRS_DET_SAVE_DW", "1") == "1"
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class View:
    """Planar NCHW fp32 view: channel stride is H*W, sample stride `ss`, optional load transform."""

    __slots__ = ("t", "off", "ss", "C", "H", "W", "xf")

    def __init__(self, t, off, ss, C, H, W, xf=None):
        self.t, self.off, self.ss, self.C, self.H, self.W, self.xf = t, off, ss, C, H, W, xf

    @property
    def p(self):
        return self.t.data_ptr() + 4 * self.off

    def xfp(self):
        if self.xf is None:
            return (None, None, None)
        return tuple(a.data_ptr() for a in self.xf)

    def chan(self, c0, c1):
        xf = None if self.xf is None else tuple(a[c0:c1] for a in self.xf)
        return View(self.t, self.off + c0 * self.H * self.W, self.ss, c1 - c0, self.H, self.W, xf)


def new_view(N, C, H, W, dev, xf=None):
    t = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
    return View(t, 0, C * H * W, C, H, W, xf)


def _finalize(partials, nblk, K, out, st):
    call("ocrs_finalize_partials", ptr(partials), nblk, K, ptr(out), st)


class _Sep:
    """One DepthwiseConv block (models.py:7-28): parameters + what backward needs."""

    def __init__(self, mod):
        self.dw, self.pw, self.bn = mod.seq[0], mod.seq[1], mod.seq[2]
        self.cin, self.cout = self.pw.in_channels, self.pw.out_channels
        # the block whose raw output IS this block's input and whose d_a this block's depthwise backward writes last
        self.producer: _Sep | None = None

    def params(self):
        return [self.dw.weight, self.pw.weight, self.bn.weight, self.bn.bias]

    def forward(self, inp: View, N, training, st, y: View | None = None, xf_dst=None, save=None):
        """`save`: dict that receives this block's backward record (keyed by id(self)), or None."""
        dev = inp.t.device
        H, W = inp.H, inp.W
        if y is None:
            y = new_view(N, self.cout, H, W, dev)
        lib = _lib.lib()
        dwo = None
        if self._tc_ok(inp, y, H, W):
            return self._forward_tc(inp, y, N, H, W, training, st, xf_dst, save)
        tma = (USE_TMA and bool(lib.ocrs_det_tma_supported(inp.p, inp.ss, y.p, y.ss, H, W))
               and bool(lib.ocrs_det_sep_channels_ok(self.cin, self.cout)))
        meta = 4.0 * N * H * W * (self.cin + self.cout)
        if tma:  # TMA-pipelined persistent kernel (csrc/det_tma.cu)
            rows = lib.ocrs_det_sep_fwd_rows(N, H, W, self.cout)
            partials = torch.empty((rows, 2, self.cout), dtype=torch.float32, device=dev) if training else None
            # training: keep the depthwise output for the 1x1 weight gradient (the step is issue-bound, not HBM-bound:
            # 4*Cin bytes/pixel of extra traffic buy back the stencil recompute in backward)
            dwo = torch.empty((N, self.cin, H, W), dtype=torch.float32, device=dev) if (save is not None and SAVE_DW) else None
            call("ocrs_det_sep_fwd", inp.p, inp.ss, N, self.cin, H, W, *inp.xfp(), ptr(self.dw.weight),
                 ptr(self.pw.weight), self.cout, y.p, y.ss, ptr(partials), ptr(dwo), st,
                 meta=meta + (4.0 * N * H * W * self.cin if dwo is not None else 0.0))
        else:
            rows = lib.ocrs_det_dwpw_partial_rows(N, H, W)
            partials = torch.empty((rows, 2, self.cout), dtype=torch.float32, device=dev) if training else None
            call("ocrs_det_dwpw_fwd", inp.p, inp.ss, N, self.cin, H, W, *inp.xfp(), ptr(self.dw.weight),
                 ptr(self.pw.weight), self.cout, y.p, y.ss, ptr(partials), st, meta=meta)
        if xf_dst is None:
            buf = torch.empty((3, self.cout), dtype=torch.float32, device=dev)
            xf_dst = (buf[0], buf[1], buf[2])
        stats = torch.empty((2, self.cout), dtype=torch.float32, device=dev)
        bn = self.bn
        call("ocrs_bn_finalize", ptr(partials), rows, self.cout, float(N * H * W), ptr(bn.weight), ptr(bn.bias),
             ptr(bn.running_mean), ptr(bn.running_var), BN_MOMENTUM, BN_EPS, int(training), 1,
             xf_dst[0].data_ptr(), xf_dst[1].data_ptr(), xf_dst[2].data_ptr(), ptr(stats[0]), ptr(stats[1]),
             ptr(bn.num_batches_tracked), st)
        y.xf = xf_dst
        if save is not None:
            save[id(self)] = (inp, y, stats, bool(training), dwo)
        return y

    # ---- levels with >= 64 channels: the 1x1 convolution and its two gradients are GEMMs (M = channels, N = pixels)
    # on the tcgen05 kernel of the recognition path (csrc/gemm_tc.cu, batched over the samples of the planar layout) ----
    def _tc_ok(self, inp: View, y: View, H, W):
        HW = H * W
        return (USE_TC_DEEP and self.cout >= 64 and self.cin >= 32 and self.cin % 32 == 0 and self.cout % 4 == 0
                and HW % 4 == 0 and HW >= 64 and y.ss % 4 == 0 and y.p % 16 == 0)

    def _forward_tc(self, inp: View, y: View, N, H, W, training, st, xf_dst, save):
        dev = inp.t.device
        lib = _lib.lib()
        HW, ci, co = H * W, self.cin, self.cout
        dwo = torch.empty((N, ci, H, W), dtype=torch.float32, device=dev)
        call("ocrs_det_dw3x3_fwd", inp.p, inp.ss, N, ci, H, W, *inp.xfp(), ptr(self.dw.weight), ptr(dwo), st,
             meta=8.0 * N * HW * ci)
        rows = lib.ocrs_gemm_tc_batched_stat_rows(HW, N)
        partials = torch.empty((rows, 2, co), dtype=torch.float32, device=dev) if training else None
        call("ocrs_gemm_tc_batched", ptr(self.pw.weight), ci, 1, co, 0, ptr(dwo), HW, 0, N * ci, ci, y.p, HW, y.ss,
             co, HW, ci, N, ptr(partials), st, meta=2.0 * N * HW * ci * co)
        if xf_dst is None:
            buf = torch.empty((3, co), dtype=torch.float32, device=dev)
            xf_dst = (buf[0], buf[1], buf[2])
        stats = torch.empty((2, co), dtype=torch.float32, device=dev)
        bn = self.bn
        call("ocrs_bn_finalize", ptr(partials), rows, co, float(N * HW), ptr(bn.weight), ptr(bn.bias),
             ptr(bn.running_mean), ptr(bn.running_var), BN_MOMENTUM, BN_EPS, int(training), 1,
             xf_dst[0].data_ptr(), xf_dst[1].data_ptr(), xf_dst[2].data_ptr(), ptr(stats[0]), ptr(stats[1]),
             ptr(bn.num_batches_tracked), st)
        y.xf = xf_dst
        if save is not None:
            save[id(self)] = (inp, y, stats, bool(training), ("tc", dwo))
        return y

    def _pw_backward_tc(self, d_a: View, y: View, k, dwo, N, H, W, st):
        """dy -> g = W^T dy and dW = sum_n dy x^T as batched tcgen05 GEMMs. Returns (g view, d_wpw)."""
        dev = y.t.device
        HW, ci, co = H * W, self.cin, self.cout
        dy = torch.empty((N, co, HW), dtype=torch.float32, device=dev)
        call("ocrs_det_dy", d_a.p, d_a.ss, y.p, y.ss, N, co, HW, *k, ptr(dy), st, meta=12.0 * N * HW * co)
        g = new_view(N, ci, H, W, dev)
        call("ocrs_gemm_tc_batched", ptr(self.pw.weight), ci, 0, co, 0, ptr(dy), HW, 0, N * co, co, g.p, HW, g.ss,
             ci, HW, co, N, None, st, meta=2.0 * N * HW * ci * co)
        ks = _wgrad_k_splits(HW, N, co, ci)
        wpart = torch.empty((N * ks, co, ci), dtype=torch.float32, device=dev)
        call("ocrs_gemm_tc_batched_splitk", ptr(dy), HW, 1, N * co, co, ptr(dwo), HW, 1, N * ci, ci, ptr(wpart), ci, co * ci,
             co, ci, HW, N, ks, st, meta=2.0 * N * HW * ci * co)
        return g, Partial(wpart, N * ks, co * ci)

    def backward(self, saved: dict, d_a: View, N, st, dx: View | None, accumulate=False, bn_pending=None):
        """d_a: gradient w.r.t. this block's activated output. Writes the gradient w.r.t. the
        block's (activated) input into `dx`; returns [d_wdw, d_wpw, d_gamma, d_beta].
        `bn_pending`: dict id(block) -> (partials, rows) of BatchNorm-backward sums already produced by the
        kernel that wrote that block's d_a; this call consumes its own entry and, when its depthwise backward is
        the final writer of the upstream block's d_a (`self.producer`), leaves that block's entry."""
        inp, y, stats, training, dwo = saved.pop(id(self))
        dev = y.t.device
        lib = _lib.lib()
        H, W, HW = y.H, y.W, y.H * y.W
        co, ci = self.cout, self.cin
        ysc, ysh, ylo = y.xfp()
        mine = bn_pending.pop(id(self), None) if bn_pending is not None else None
        if mine is not None:
            part, rows = mine
        else:
            rows = lib.ocrs_reduce_rows(N, HW)
            part = torch.empty((rows, 2, co), dtype=torch.float32, device=dev)
            call("ocrs_bnrelu_bwd_reduce", d_a.p, d_a.ss, y.p, y.ss, N, co, HW, ysc, ysh, ylo, ptr(stats[0]),
                 ptr(stats[1]), ptr(part), st, meta=4.0 * N * HW * 2 * co)
        coef = torch.empty((5, co), dtype=torch.float32, device=dev)  # dgamma, dbeta, k1, k2, k3
        call("ocrs_bn_bwd_finalize", ptr(part), rows, co, float(N * HW), ptr(self.bn.weight), ptr(stats[0]),
             ptr(stats[1]), ptr(coef[0]), ptr(coef[1]), ptr(coef[2]), ptr(coef[3]), ptr(coef[4]), int(training), st)
        k = (ysc, ysh, ylo, ptr(coef[2]), ptr(coef[3]), ptr(coef[4]))
        if isinstance(dwo, tuple):  # levels with >= 64 channels: batched tcgen05 GEMMs
            g, d_wpw = self._pw_backward_tc(d_a, y, k, dwo[1], N, H, W, st)
        else:
            g = new_view(N, ci, H, W, dev)
            fuse_g = dwo is not None and co <= 16 and FUSE_PWT  # the weight-gradient kernel also emits g (csrc/det_tma.cu)
            if dwo is not None and co == 32 and ci % 16 == 0 and FUSE_PWT and HW % 4 == 0:
                # 32 output channels: one CTA holds them all, so g comes out of the same staged tile here too
                workers = lib.ocrs_det_pw_wgrad_saved32_workers(N, HW, ci)
                wpart = torch.empty((workers, co, ci), dtype=torch.float32, device=dev)
                call("ocrs_det_pw_wgrad_saved32", d_a.p, d_a.ss, y.p, y.ss, N, HW, *k, ptr(dwo), ci, ptr(wpart),
                     ptr(self.pw.weight), g.p, g.ss, st, meta=4.0 * N * HW * (2 * co + 2 * ci))
                d_wpw = Partial(wpart, workers, co * ci)
                dwo = None
                return self._dw_backward(saved, inp, g, dx, d_wpw, coef, N, H, W, st, accumulate, bn_pending)
            if not fuse_g:
                call("ocrs_det_pwT_bwd", d_a.p, d_a.ss, y.p, y.ss, N, co, HW, *k, ptr(self.pw.weight), ci, g.p, g.ss, st,
                     meta=4.0 * N * HW * (2 * co + ci))
            if dwo is not None:
                workers = lib.ocrs_det_pw_wgrad_saved_workers(N, HW, co, ci)
                wpart = torch.empty((workers, co, ci), dtype=torch.float32, device=dev)
                call("ocrs_det_pw_wgrad_saved", d_a.p, d_a.ss, y.p, y.ss, N, co, HW, *k, ptr(dwo), ci, ptr(wpart),
                     ptr(self.pw.weight) if fuse_g else None, g.p if fuse_g else None, g.ss, st,
                     meta=4.0 * N * HW * (2 * co + ci + (ci if fuse_g else 0)))
            elif USE_TMA and lib.ocrs_det_tma_supported(inp.p, inp.ss, inp.p, inp.ss, H, W):
                workers = lib.ocrs_det_sep_pw_wgrad_workers(N, H, W, co, ci)
                wpart = torch.empty((workers, co, ci), dtype=torch.float32, device=dev)
                call("ocrs_det_sep_pw_wgrad", d_a.p, d_a.ss, y.p, y.ss, N, co, H, W, *k, inp.p, inp.ss, ci, *inp.xfp(),
                     ptr(self.dw.weight), ptr(wpart), st, meta=4.0 * N * HW * (2 * co + ci))
            else:
                workers = lib.ocrs_det_pw_wgrad_workers(N, H, W)
                wpart = torch.empty((workers, co, ci), dtype=torch.float32, device=dev)
                call("ocrs_det_pw_wgrad", d_a.p, d_a.ss, y.p, y.ss, N, co, H, W, *k, inp.p, inp.ss, ci, *inp.xfp(),
                     ptr(self.dw.weight), ptr(wpart), st, meta=4.0 * N * HW * (2 * co + ci))
            d_wpw = Partial(wpart, workers, co * ci)
        dwo = None
        return self._dw_backward(saved, inp, g, dx, d_wpw, coef, N, H, W, st, accumulate, bn_pending)

    def _dw_backward(self, saved, inp, g, dx, d_wpw, coef, N, H, W, st, accumulate, bn_pending):
        """Depthwise 3x3 backward of the block from g = d(depthwise output): dx, the depthwise weight gradient and, when this
        block is the final writer of the upstream block's d_a, that block's BatchNorm-backward sums."""
        dev = g.t.device
        lib = _lib.lib()
        HW, ci = H * W, self.cin
        if dx is None:
            dx = new_view(N, ci, H, W, dev)
        if USE_TMA and lib.ocrs_det_tma_supported(g.p, g.ss, inp.p, inp.ss, H, W) and lib.ocrs_det_tma_supported(dx.p, dx.ss, dx.p, dx.ss, H, W):
            drows = lib.ocrs_det_sep_dw_bwd_rows(N, H, W, ci)
            dpart = torch.empty((drows, ci, 9), dtype=torch.float32, device=dev)
            up = self.producer if (FUSE_BN_REDUCE and bn_pending is not None and self.producer is not None and id(self.producer) in saved) else None
            if up is not None:
                up_stats = saved[id(up)][2]
                bnp = torch.empty((drows, 2, ci), dtype=torch.float32, device=dev)
                call("ocrs_det_sep_dw_bwd", g.p, g.ss, inp.p, inp.ss, N, ci, H, W, *inp.xfp(), ptr(self.dw.weight), dx.p,
                     dx.ss, int(accumulate), ptr(dpart), ptr(up_stats[0]), ptr(up_stats[1]), ptr(bnp), st,
                     meta=4.0 * N * HW * (3 + int(accumulate)) * ci)
                bn_pending[id(up)] = (bnp, drows)
            else:
                call("ocrs_det_sep_dw_bwd", g.p, g.ss, inp.p, inp.ss, N, ci, H, W, *inp.xfp(), ptr(self.dw.weight), dx.p,
                     dx.ss, int(accumulate), ptr(dpart), None, None, None, st, meta=4.0 * N * HW * (3 + int(accumulate)) * ci)
        else:
            drows = lib.ocrs_det_dw_bwd_rows(N, H, W)
            dpart = torch.empty((drows, ci, 9), dtype=torch.float32, device=dev)
            call("ocrs_det_dw_bwd", g.p, g.ss, inp.p, inp.ss, N, ci, H, W, *inp.xfp(), ptr(self.dw.weight), dx.p,
                 dx.ss, int(accumulate), ptr(dpart), st, meta=4.0 * N * HW * 3 * ci)
        # weight gradients stay partial rows; grads.deliver reduces all of the step's in one launch
        return [Partial(dpart, drows, ci * 9), d_wpw, coef[0], coef[1]], dx


class _Plan:
    """Static description of the network built from the module tree (parameter order included)."""

    def __init__(self, model):
        self.model = model
        d = model.depth_scale
        self.levels = len(d) - 1
        self.in_conv = [_Sep(model.in_conv.seq[0]), _Sep(model.in_conv.seq[1])]
        self.down = [[_Sep(m.seq[0].seq[0]), _Sep(m.seq[0].seq[1])] for m in model.down]
        self.upT = [m.up for m in model.up]
        self.contract = [[_Sep(m.contract.seq[0]), _Sep(m.contract.seq[1])] for m in model.up]
        self.out = model.out_conv[0]
        self.depth = d
        self.in_conv[1].producer = self.in_conv[0]
        self.down[0][0].producer = self.in_conv[1]  # skip connection: this (accumulating) backward runs after Up's
        for pair in self.down + self.contract:
            pair[1].producer = pair[0]

    def all_params(self):
        ps = []
        for s in self.in_conv:
            ps += s.params()
        for pair in self.down:
            for s in pair:
                ps += s.params()
        for i in range(self.levels):
            ps += [self.upT[i].weight, self.upT[i].bias]
            for s in self.contract[i]:
                ps += s.params()
        ps += [self.out.weight, self.out.bias]
        return ps


def _wgrad_k_splits(HW, batch, M, Nn):
    """K slices per sample for a batched weight-gradient GEMM (K = pixels of one sample): enough tiles to fill the GPU
    twice, at least 512 pixels per slice."""
    bn = 32 if Nn <= 32 else 64 if Nn <= 64 else 128
    tiles = batch * ((M + 127) // 128) * ((Nn + bn - 1) // bn)
    want = max(1, min(296 // max(tiles, 1), HW // 512))
    return _lib.lib().ocrs_gemm_tc_batched_splits(HW, want)


def _convt_tc_ok(t, up: View, cout):
    """ConvTranspose2d (models.py:76-78) as batched tcgen05 GEMMs: Z[n] = W^T x[n] with W read as [Cin][9*Cout]."""
    HW = up.H * up.W
    return (USE_TC_DEEP and up.C >= 64 and up.C % 32 == 0 and cout % 32 == 0 and HW % 4 == 0 and HW >= 64
            and t.weight.data_ptr() % 16 == 0)


def _convt_forward_tc(t, up: View, N, cout, out: View, keep, st):
    dev = up.t.device
    ci, HW = up.C, up.H * up.W
    xa = torch.empty((N, ci, HW), dtype=torch.float32, device=dev)
    call("ocrs_det_activate", up.p, up.ss, N, ci, HW, *up.xfp(), ptr(xa), st, meta=8.0 * N * HW * ci)
    z = torch.empty((N, 9 * cout, HW), dtype=torch.float32, device=dev)
    call("ocrs_gemm_tc_batched", ptr(t.weight), 9 * cout, 0, ci, 0, ptr(xa), HW, 0, N * ci, ci, ptr(z), HW,
         9 * cout * HW, 9 * cout, HW, ci, N, None, st, meta=2.0 * N * HW * ci * 9 * cout)
    call("ocrs_det_convt_col2im", ptr(z), N, cout, up.H, up.W, ptr(t.bias), out.p, out.ss, out.H, out.W, st,
         meta=4.0 * N * (9 * cout * HW + cout * out.H * out.W))
    return xa if keep else None


def _convt_backward_tc(t, xa, dlo: View, N, ci, Hin, Win, st):
    """Returns (d_up view = gradient w.r.t. the activated input, dW)."""
    dev = xa.device
    c, HW = dlo.C, Hin * Win
    dcol = torch.empty((N, 9 * c, HW), dtype=torch.float32, device=dev)
    call("ocrs_det_convt_im2col", dlo.p, dlo.ss, N, c, dlo.H, dlo.W, Hin, Win, ptr(dcol), st,
         meta=4.0 * N * (9 * c * HW + c * dlo.H * dlo.W))
    d_up = new_view(N, ci, Hin, Win, dev)
    call("ocrs_gemm_tc_batched", ptr(t.weight), 9 * c, 1, ci, 0, ptr(dcol), HW, 0, N * 9 * c, 9 * c, d_up.p, HW, d_up.ss,
         ci, HW, 9 * c, N, None, st, meta=2.0 * N * HW * ci * 9 * c)
    ks = _wgrad_k_splits(HW, N, ci, 9 * c)
    wpart = torch.empty((N * ks, ci, 9 * c), dtype=torch.float32, device=dev)
    call("ocrs_gemm_tc_batched_splitk", ptr(xa), HW, 1, N * ci, ci, ptr(dcol), HW, 1, N * 9 * c, 9 * c, ptr(wpart), 9 * c,
         ci * 9 * c, ci, 9 * c, HW, N, ks, st, meta=2.0 * N * HW * ci * 9 * c)
    return d_up, Partial(wpart, N * ks, ci * 9 * c)


_IDENT: dict = {}


def _identity_xf(C, dev):
    """(scale, shift, lo) = (1, 0, -inf): constants, built once per (channels, device)."""
    key = (C, dev)
    xf = _IDENT.get(key)
    if xf is None:
        buf = torch.empty((3, C), dtype=torch.float32, device=dev)
        buf[0].fill_(1.0)
        buf[1].zero_()
        buf[2].fill_(NEG_INF)
        xf = _IDENT[key] = (buf[0], buf[1], buf[2])
    return xf


class _DetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: _Plan, x, *params):
        model = plan.model
        training = model.training
        dev = x.device
        N, _, H, W = x.shape
        d = plan.depth
        L = plan.levels
        st = _lib.stream_ptr(dev)
        save = {} if any(ctx.needs_input_grad) else None
        xin = View(x, 0, H * W, 1, H, W)
        # resolution of each level: level 0 = input size, level i = floor(level i-1 / 2)
        hs, ws = [H], [W]
        for i in range(L):
            hs.append(hs[-1] // 2)
            ws.append(ws[-1] // 2)
        if hs[L] < 1 or ws[L] < 1:
            raise RuntimeError(f"DetectionModel needs inputs of at least {2**L}x{2**L}, got {H}x{W}")
        # concat buffers of the six Up stages: [ConvT output | skip]
        cat = [torch.empty((N, 2 * d[i], hs[i], ws[i]), dtype=torch.float32, device=dev) for i in range(L)]
        # their load transforms: identity for the ConvT half, the skip producer's BatchNorm+ReLU (written by its
        # bn_finalize) for the other half - per forward pass, one copy of the constant identity rows for all six
        ident = _identity_xf(2 * sum(d[:L]), dev)
        xfbuf = torch.stack(ident)
        coff = [2 * sum(d[:i]) for i in range(L + 1)]
        catxf = [tuple(xfbuf[r, coff[i] : coff[i + 1]] for r in range(3)) for i in range(L)]

        def cat_view(i, lo, hi):
            xf = tuple(a[lo:hi] for a in catxf[i])
            return View(cat[i], lo * hs[i] * ws[i], 2 * d[i] * hs[i] * ws[i], hi - lo, hs[i], ws[i], xf)

        with torch.cuda.device(dev):
            a = plan.in_conv[0].forward(xin, N, training, st, save=save)
            # in_conv output (raw + its BN transform) lives in the skip half of cat[0]
            skip0 = cat_view(0, d[0], 2 * d[0])
            xfull = plan.in_conv[1].forward(a, N, training, st, y=skip0, xf_dst=skip0.xf, save=save)
            prev = xfull
            pooled_src = []
            for i in range(L):
                a = plan.down[i][0].forward(prev, N, training, st, save=save)
                b = plan.down[i][1].forward(a, N, training, st, save=save)
                if i + 1 < L:
                    dst = cat_view(i + 1, d[i + 1], 2 * d[i + 1])
                else:
                    dst = new_view(N, d[L], hs[L], ws[L], dev, _identity_xf(d[L], dev))
                call("ocrs_det_pool2_fwd", b.p, b.ss, N, b.C, b.H, b.W, *b.xfp(), dst.p, dst.ss, st)
                pooled_src.append(b)
                prev = dst
            up = prev
            up_inputs = []
            for i in reversed(range(L)):
                t = plan.upT[i]
                lo_half = cat_view(i, 0, d[i])
                xa = None
                if _convt_tc_ok(t, up, d[i]):
                    xa = _convt_forward_tc(t, up, N, d[i], lo_half, save is not None, st)
                else:
                    call("ocrs_det_convt_fwd", up.p, up.ss, N, up.C, up.H, up.W, *up.xfp(), ptr(t.weight), ptr(t.bias),
                         d[i], lo_half.p, lo_half.ss, hs[i], ws[i], st)
                up_inputs.append((i, (up, xa)))
                full = cat_view(i, 0, 2 * d[i])
                a = plan.contract[i][0].forward(full, N, training, st, save=save)
                up = plan.contract[i][1].forward(a, N, training, st, save=save)
            prob = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
            call("ocrs_det_outconv_fwd", up.p, up.ss, N, d[0], H, W, *up.xfp(), ptr(plan.out.weight),
                 ptr(plan.out.bias), ptr(prob), st)
        if save is not None:
            ctx.recs = save
            ctx.plan = plan
            ctx.geom = (N, H, W, hs, ws)
            ctx.acts = (xin, cat, catxf, pooled_src, dict(up_inputs), up, prob)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        plan: _Plan = ctx.plan
        N, H, W, hs, ws = ctx.geom
        xin, cat, catxf, pooled_src, up_inputs, last, prob = ctx.acts
        d, L = plan.depth, plan.levels
        dev = prob.device
        lib = _lib.lib()
        st = _lib.stream_ptr(dev)
        dprob = dprob.contiguous().float()
        grads: dict = {}
        recs = ctx.recs
        pend: dict = {}

        def put(tensors, gs):
            for t, g in zip(tensors, gs):
                grads[id(t)] = g

        with torch.cuda.device(dev):
            # out_conv + sigmoid (+ the BatchNorm-backward sums of the block feeding it)
            d_a = new_view(N, d[0], H, W, dev)
            last_blk = plan.contract[0][1]
            if FUSE_BN_REDUCE and d[0] == 8 and (H * W) % 4 == 0 and id(last_blk) in recs and last.xf is not None:
                blocks = lib.ocrs_det_outconv8_bwd_blocks()
                part = torch.empty((blocks, d[0] + 1), dtype=torch.float32, device=dev)
                bnp = torch.empty((blocks, 2, d[0]), dtype=torch.float32, device=dev)
                st_last = recs[id(last_blk)][2]
                call("ocrs_det_outconv8_bwd_bn", ptr(dprob), ptr(prob), last.p, last.ss, N, H * W, *last.xfp(),
                     ptr(plan.out.weight), d_a.p, d_a.ss, ptr(st_last[0]), ptr(st_last[1]), ptr(part), ptr(bnp), st,
                     meta=4.0 * N * H * W * (2 + 2 * d[0]))
                pend[id(last_blk)] = (bnp, blocks)
                rows = blocks
            else:
                rows = lib.ocrs_det_outconv_bwd_rows(N, H, W)
                part = torch.empty((rows, d[0] + 1), dtype=torch.float32, device=dev)
                call("ocrs_det_outconv_bwd", ptr(dprob), ptr(prob), last.p, last.ss, N, d[0], H, W, *last.xfp(),
                     ptr(plan.out.weight), d_a.p, d_a.ss, ptr(part), st)
            put([plan.out.weight, plan.out.bias],
                [Partial(part, rows, d[0], ld=d[0] + 1), Partial(part, rows, 1, ld=d[0] + 1, off=d[0])])

            dcat = [None] * L
            # up path, in reverse of forward order: up[0] first
            for i in range(L):
                c = d[i]
                gB, d_mid = plan.contract[i][1].backward(recs, d_a, N, st, None, bn_pending=pend)
                put(plan.contract[i][1].params(), gB)
                dcat[i] = new_view(N, 2 * c, hs[i], ws[i], dev)
                gA, _ = plan.contract[i][0].backward(recs, d_mid, N, st, dcat[i], bn_pending=pend)
                put(plan.contract[i][0].params(), gA)
                # ConvTranspose2d
                t = plan.upT[i]
                up_in, xa = up_inputs[i]
                dlo = dcat[i].chan(0, c)
                if xa is not None:  # >= 64 input channels: im2col + two batched tcgen05 GEMMs
                    d_up, dw = _convt_backward_tc(t, xa, dlo, N, up_in.C, up_in.H, up_in.W, st)
                    xa = None
                else:
                    d_up = new_view(N, up_in.C, up_in.H, up_in.W, dev)
                    call("ocrs_det_convt_bwd_data", dlo.p, dlo.ss, N, c, hs[i], ws[i], ptr(t.weight), up_in.C, up_in.H,
                         up_in.W, d_up.p, d_up.ss, st)
                    if USE_TMA and lib.ocrs_det_convt_wgrad_staged_ok(up_in.p, up_in.ss, up_in.H, up_in.W, dlo.p, dlo.ss, hs[i], ws[i]):
                        workers = lib.ocrs_det_convt_wgrad_staged_workers(N, up_in.H, up_in.W, up_in.C, c)
                        wpart = torch.empty((workers,) + tuple(t.weight.shape), dtype=torch.float32, device=dev)
                        call("ocrs_det_convt_wgrad_staged", up_in.p, up_in.ss, N, up_in.C, up_in.H, up_in.W, *up_in.xfp(),
                             dlo.p, dlo.ss, c, hs[i], ws[i], ptr(wpart), st)
                    else:
                        workers = lib.ocrs_det_convt_wgrad_workers(N, up_in.H, up_in.W)
                        wpart = torch.empty((workers,) + tuple(t.weight.shape), dtype=torch.float32, device=dev)
                        call("ocrs_det_convt_wgrad", up_in.p, up_in.ss, N, up_in.C, up_in.H, up_in.W, *up_in.xfp(), dlo.p,
                             dlo.ss, c, hs[i], ws[i], ptr(wpart), st)
                    dw = Partial(wpart, workers, t.weight.numel())
                brows = lib.ocrs_reduce_rows(N, hs[i] * ws[i])
                bpart = torch.empty((brows, c), dtype=torch.float32, device=dev)
                call("ocrs_plane_sum", dlo.p, dlo.ss, N, c, hs[i] * ws[i], ptr(bpart), st)
                put([t.weight, t.bias], [dw, Partial(bpart, brows, c)])
                d_a = d_up  # gradient w.r.t. the activated input of this ConvT
            # d_a now = gradient w.r.t. x_down[L-1] (pooled, activated)
            d_pooled = d_a
            for i in reversed(range(L)):
                b = pooled_src[i]
                d_full = new_view(N, b.C, b.H, b.W, dev)
                blk = plan.down[i][1]
                if FUSE_BN_REDUCE and id(blk) in recs and b.xf is not None:
                    prow = lib.ocrs_det_pool2_bwd_bn_rows(N, b.H, b.W)
                    bnp = torch.empty((prow, 2, b.C), dtype=torch.float32, device=dev)
                    st_b = recs[id(blk)][2]
                    call("ocrs_det_pool2_bwd_bn", b.p, b.ss, N, b.C, b.H, b.W, *b.xfp(), d_pooled.p, d_pooled.ss,
                         d_full.p, d_full.ss, ptr(st_b[0]), ptr(st_b[1]), ptr(bnp), st)
                    pend[id(blk)] = (bnp, prow)
                else:
                    call("ocrs_det_pool2_bwd", b.p, b.ss, N, b.C, b.H, b.W, *b.xfp(), d_pooled.p, d_pooled.ss,
                         d_full.p, d_full.ss, st)
                gB, d_mid = plan.down[i][1].backward(recs, d_full, N, st, None, bn_pending=pend)
                put(plan.down[i][1].params(), gB)
                # input of down[i] is the skip half of cat[i]; its gradient already holds the Up-path part
                dskip = dcat[i].chan(d[i], 2 * d[i])
                gA, _ = plan.down[i][0].backward(recs, d_mid, N, st, dskip, accumulate=True, bn_pending=pend)
                put(plan.down[i][0].params(), gA)
                d_pooled = dskip
            # in_conv: d_pooled is now the gradient w.r.t. the activated in_conv output
            gB, d_mid = plan.in_conv[1].backward(recs, d_pooled, N, st, None, bn_pending=pend)
            put(plan.in_conv[1].params(), gB)
            gA, dx = plan.in_conv[0].backward(recs, d_mid, N, st, None, bn_pending=pend)
            put(plan.in_conv[0].params(), gA)
            param_grads = deliver(plan.all_params(), grads, st)
        dxt = dx.t if ctx.needs_input_grad[1] else None
        out = [None, dxt] + param_grads
        ctx.acts = None
        return tuple(out)


def detection_forward(model, x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("ocrs_models_b200.DetectionModel has no CPU path: input must be a CUDA tensor")
    if x.dim() != 4 or x.shape[1] != 1:
        raise RuntimeError(f"expected (N, 1, H, W) input, got {tuple(x.shape)}")
    plan = model.__dict__.get("_plan")
    if plan is None:
        plan = _Plan(model)
        model.__dict__["_plan"] = plan
    if model.__dict__.get("_ocrs_checked") != x.device:
        _lib.check_module_tensors(model, x.device, "DetectionModel")
        model.__dict__["_ocrs_checked"] = x.device
    x = x.float().contiguous()
    return _DetFunction.apply(plan, x, *plan.all_params())
