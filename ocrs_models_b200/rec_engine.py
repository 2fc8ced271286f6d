"""Execution plan of the recognition CRNN on the sm_100a kernels (csrc/rec.cu, gemm.cu).

Follows reference ``RecognitionModel.forward`` (ocrs_models/models.py:253-268): conv stack
(models.py:179-243) in NHWC, permute to (W', N, C) for free in the last pooling kernel's store,
2-layer bidirectional GRU in fp32 (the reference also forces fp32 there, models.py:264-266),
Linear + LogSoftmax. One ``torch.autograd.Function`` for the whole model; its backward runs the
hand-written BPTT / dgrad / wgrad kernels.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr
from .grads import Partial, deliver

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
TARGET_BLOCKS = 2 * 148
# "tc": tcgen05 3xTF32 tensor-core GEMM (csrc/gemm_tc.cu) wherever TMA alignment allows, else the
# fp32 CUDA-core GEMM (csrc/gemm.cu). OCRS_GEMM=simt forces the latter (A/B testing).
GEMM_BACKEND = os.environ.get("OCRS_GEMM", "tc")
# OCRS_EXACT_FWD=1 runs the forward convolutions on the fp32-FMA GEMM instead of the tensor cores.
EXACT_FWD = os.environ.get("OCRS_EXACT_FWD", "0") == "1"
# Implicit-GEMM 3x3 convolutions (im2col folded into the TMA coordinates); OCRS_IMPLICIT=0 goes back to
# an explicit im2col buffer + GEMM.
IMPLICIT = os.environ.get("OCRS_IMPLICIT", "1") == "1"
# Weight operands of the tensor-core GEMMs are split into TF32 hi/lo once per use in HBM (ocrs_split_tf32) and
# both halves arrive by TMA, so the converter warps only touch the activation operand. OCRS_PRESPLIT=0 converts
# the weights inside the GEMM like the activations.
PRESPLIT = os.environ.get("OCRS_PRESPLIT", "1") == "1"


def set_precision(mode: str) -> str:
    """Numerics mode of the recognition path's tensor-core GEMMs: "parity" (default: 3xTF32 + four TMEM accumulators,
    fp32-class accuracy, what the 1e-3 parity claim and every headline number refer to) or "tf32" (labelled fast
    mode: one plain TF32 product per k-step - the numerics class of what reference train_rec.py:118 runs on a GPU under
    autocast, NOT parity numerics; GRU recurrence, BatchNorm statistics, CTC and the optimiser stay fp32).
    Returns the previous mode. Also selectable with OCRS_PRECISION=tf32."""
    global PRESPLIT, _PRECISION
    if mode not in ("parity", "tf32"):
        raise ValueError("precision mode must be 'parity' or 'tf32'")
    prev = _PRECISION
    _lib.lib().ocrs_gemm_tc_set_fast(int(mode == "tf32"))
    PRESPLIT = mode == "parity" and os.environ.get("OCRS_PRESPLIT", "1") == "1"
    _PRECISION = mode
    return prev


_PRECISION = "parity"


class Split:
    """A weight matrix split for the 3xTF32 GEMM: `hi` (TF32-exact) and `lo` (remainder), same layout."""

    __slots__ = ("hi", "lo", "src")

    def __init__(self, w, st, halves=None):
        self.src = w  # the unsplit matrix in the same layout (None if there is none), for GEMMs that cannot use the halves
        if halves is not None:  # already produced (ocrs_weight_prep)
            self.hi, self.lo = halves
            return
        w = w.detach().contiguous()
        self.hi = torch.empty_like(w)
        self.lo = torch.empty_like(w)
        call("ocrs_split_tf32", ptr(w), ptr(self.hi), ptr(self.lo), w.numel(), st)


def _maybe_split(w, st):
    return Split(w, st) if (PRESPLIT and GEMM_BACKEND == "tc") else w


def conv3x3(x, N, H, W, cin, wp, cout, st, bias=None, relu=False, stats=None):
    """3x3 / pad 1 convolution of an NHWC tensor on the tensor cores without an im2col buffer.
    wp: [cout, (ky, kx, ci)]. Returns [N*H*W, cout]."""
    out = _empty((N * H * W, cout), x.device)
    if isinstance(wp, torch.Tensor):
        wp = _maybe_split(wp, st)
    if isinstance(wp, Split):
        call("ocrs_conv3x3_tc_presplit", ptr(x), N, H, W, cin, ptr(wp.hi), ptr(wp.lo), cout, ptr(out), cout, ptr(bias),
             int(relu), ptr(stats), st, meta=2.0 * N * H * W * cout * 9 * cin)
    else:
        call("ocrs_conv3x3_tc", ptr(x), N, H, W, cin, ptr(wp), cout, ptr(out), cout, ptr(bias), int(relu), ptr(stats), st,
             meta=2.0 * N * H * W * cout * 9 * cin)
    return out


def conv3x3_wgrad(dy, x, N, H, W, cin, cout, st):
    """[cout, (ky, kx, ci)] weight gradient of conv3x3 from dy [N*H*W, cout] and the NHWC input x, as split-K partial rows."""
    lib = _lib.lib()
    K = N * H * W
    tiles = ((9 * cin + 127) // 128) * ((cout + 127) // 128)
    want = max(1, min(TARGET_BLOCKS // tiles, K // 256))
    splits = lib.ocrs_gemm_tc_splits(K, want)
    part = _empty((splits, cout, 9 * cin), x.device)
    call("ocrs_conv3x3_wgrad_tc", ptr(dy), ptr(x), N, H, W, cin, cout, ptr(part), splits, st,
         meta=2.0 * K * cout * 9 * cin)
    return Partial(part, splits, cout * 9 * cin, conv=(cin, 9))  # split-K rows, reduced + re-laid out by grads.deliver


def _implicit_ok(cin, kh):
    return IMPLICIT and GEMM_BACKEND == "tc" and not EXACT_FWD and kh == 3 and cin % 32 == 0


def _empty(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


def gemm(A, lda, a_kmajor, B, ldb, b_kmajor, M, N, K, st, out=None, ldc=None, bias=None, relu=False,
         accumulate=False, stats=None, split_ok=False, exact=False, b_weight=False, lazy=None):
    """C[M,N] = op(A) op(B). A/B are tensors or raw pointers. With `split_ok` the reduction is split
    over K when the output has too few tiles to fill the GPU. `exact` forces the fp32-FMA kernel:
    outputs that feed ReLU / max-pool decisions must be accurate to ~1e-6, because a perturbation d
    of a pre-activation flips a fraction ~d of the gates and moves gradients by ~sqrt(d). Kept as
    an A/B switch: the 4-accumulator 3xTF32 tensor-core kernel reaches the same accuracy.
    `lazy` (a dict of Partial keyword arguments, with `split_ok`): the result is a parameter gradient - return the split-K
    rows as a grads.Partial instead of reducing them here."""
    dev = out.device if out is not None else (A.device if isinstance(A, torch.Tensor) else None)
    pa = A.data_ptr() if isinstance(A, torch.Tensor) else A
    b_lo = b_src = None
    if isinstance(B, Split):
        b_lo, b_src, B = B.lo, B.src, B.hi
    pb = B.data_ptr() if isinstance(B, torch.Tensor) else B
    if out is None:
        out = _empty((M, N), dev)
    if ldc is None:
        ldc = N
    lib = _lib.lib()
    tc = GEMM_BACKEND == "tc" and not exact and bool(lib.ocrs_gemm_tc_supported(pa, lda, pb, ldb))
    if b_lo is not None and not tc and b_src is not None:  # e.g. lda = 97 (the linear head's data gradient): fp32 GEMM on the unsplit matrix
        pb, b_lo = b_src.data_ptr(), None
    fn = "ocrs_gemm_tc" if tc else "ocrs_gemm"
    splits = 1
    if split_ok:
        bn = (32 if N <= 32 else 64 if N <= 64 else 128) if tc else 128
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
        want = max(1, min(TARGET_BLOCKS // tiles, K // 256))
        splits = (lib.ocrs_gemm_tc_splits if tc else lib.ocrs_gemm_splits)(K, want)
    if b_weight and tc and PRESPLIT and splits == 1 and b_lo is None and isinstance(B, torch.Tensor):
        sp = Split(B, st)  # B is a weight matrix: split it once in HBM instead of per tile in the GEMM
        pb, b_lo = sp.hi.data_ptr(), sp.lo
    if b_lo is not None and not tc:
        raise RuntimeError("pre-split weights need the tensor-core GEMM")
    if splits == 1 and b_lo is not None:
        call("ocrs_gemm_tc_presplit", pa, lda, int(a_kmajor), pb, ptr(b_lo), ldb, int(b_kmajor), ptr(out), ldc, M, N, K,
             ptr(bias), int(relu), int(accumulate), ptr(stats), 1, st, meta=2.0 * M * N * K)
    elif splits == 1:
        call(fn, pa, lda, int(a_kmajor), pb, ldb, int(b_kmajor), ptr(out), ldc, M, N, K, ptr(bias),
             int(relu), int(accumulate), ptr(stats), 1, st, meta=2.0 * M * N * K)
    else:
        assert ldc == N and bias is None and not relu and not accumulate and stats is None
        part = _empty((splits, M, N), out.device)
        call(fn, pa, lda, int(a_kmajor), pb, ldb, int(b_kmajor), ptr(part), N, M, N, K, None, 0, 0, None,
             splits, st, meta=2.0 * M * N * K)
        if lazy is not None:
            return Partial(part, splits, M * N, **lazy)
        call("ocrs_finalize_partials", ptr(part), splits, M * N, ptr(out), st)
    if lazy is not None:
        assert ldc == N
        return Partial(out, 1, M * N, **lazy)
    return out


def colsum(A, lda, M, N, st, dev):
    lib = _lib.lib()
    rows = lib.ocrs_colsum_rows(M)
    part = _empty((rows, N), dev)
    pa = A.data_ptr() if isinstance(A, torch.Tensor) else A
    call("ocrs_colsum", pa, lda, M, N, ptr(part), st)
    return Partial(part, rows, N)


def im2col(x, N, H, W, C, kh, kw, ph, pw, st):
    Ho, Wo = H + 2 * ph - kh + 1, W + 2 * pw - kw + 1
    col = _empty((N * Ho * Wo, kh * kw * C), x.device)
    call("ocrs_im2col_nhwc", ptr(x), N, H, W, C, kh, kw, ph, pw, Ho, Wo, ptr(col), st)
    return col, Ho, Wo


def _w_fwd(w):
    """[Cout, Cin, kh, kw] -> [Cout, (ky, kx, ci)] matching the im2col column order (layout reference for the tests; the
    step builds it with ocrs_weight_prep)."""
    return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _w_dgrad(w):
    """[Cout, Cin, kh, kw] -> [Cin, (ky', kx', co)] with the kernel flipped (correlation of dY)."""
    return w.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()


def prepare_weights(model, st, dev, backward: bool):
    """Every per-step weight operand of the GEMMs in ONE launch (csrc/glue.cu ocrs_weight_prep): convolution weights
    re-laid out for the implicit GEMM (forward [Cout][(ky,kx,ci)]; flipped [Cin][(ky,kx,co)] for the data gradient),
    all GEMM weights split into TF32 hi/lo, W_hh transposed for the GRU backward. Returns dicts keyed by id(parameter):
    fwd / dg (convolutions), w (GRU input projections, linear head), whhT."""
    import ctypes

    split = PRESPLIT and GEMM_BACKEND == "tc"
    split_conv = split and not EXACT_FWD
    ent, prep = [], dict(fwd={}, dg={}, w={}, whhT={})

    def pair(shape, want):
        hi = _empty(shape, dev)
        return hi, (_empty(shape, dev) if want else None)

    for name in ("3", "7", "9", "13", "15", "19"):
        w = model.conv[name].weight
        co, ci, kh, kw = w.shape
        hi, lo = pair((co, kh * kw * ci), split_conv)
        dhi, dlo = pair((ci, kh * kw * co), split_conv) if backward else (None, None)
        ent.append((ptr(w), ptr(hi), ptr(lo), ptr(dhi), ptr(dlo), w.numel(), 1, co, ci, kh, kw))
        prep["fwd"][id(w)] = Split(None, st, (hi, lo)) if lo is not None else hi
        if backward:
            prep["dg"][id(w)] = Split(None, st, (dhi, dlo)) if dlo is not None else dhi
    if split:
        ws = [getattr(model.gru, f"weight_ih_l{l}{sfx}") for l in range(2) for sfx in ("", "_reverse")] + [model.output[0].weight]
        for w in ws:
            hi, lo = pair(tuple(w.shape), True)
            ent.append((ptr(w), ptr(hi), ptr(lo), None, None, w.numel(), 0, 0, 0, 0, 0))
            prep["w"][id(w)] = Split(w, st, (hi, lo))
    if backward:
        for l in range(2):
            for sfx in ("", "_reverse"):
                w = getattr(model.gru, f"weight_hh_l{l}{sfx}")
                t_ = _empty((w.shape[1], w.shape[0]), dev)
                ent.append((ptr(w), ptr(t_), None, None, None, w.numel(), 2, w.shape[0], w.shape[1], 0, 0))
                prep["whhT"][id(w)] = t_
    n = len(ent)
    cols = list(zip(*ent))
    ptrs = [(ctypes.c_void_p * n)(*c) for c in cols[:5]]
    ints = [(ctypes.c_int * n)(*c) for c in cols[5:]]
    call("ocrs_weight_prep", *ptrs, *ints, n, st)
    return prep


class _BNState:
    def __init__(self, bn, dev):
        self.bn = bn
        self.buf = _empty((5, bn.num_features), dev)  # scale, shift, lo(unused), mean, invstd

    @property
    def scale(self):
        return self.buf[0]

    @property
    def shift(self):
        return self.buf[1]

    @property
    def mean(self):
        return self.buf[3]

    @property
    def invstd(self):
        return self.buf[4]


def _bn_finalize(bn, stats, rows, count, training, relu, st, dev):
    s = _BNState(bn, dev)
    C = bn.num_features
    call("ocrs_bn_finalize", ptr(stats), rows, C, float(count), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
         ptr(bn.running_var), BN_MOMENTUM, BN_EPS, int(training), int(relu), ptr(s.buf[0]), ptr(s.buf[1]),
         ptr(s.buf[2]), ptr(s.buf[3]), ptr(s.buf[4]), ptr(bn.num_batches_tracked), st)
    return s


class _RecFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        dev = x.device
        training = model.training
        N, _, H, W = x.shape
        st = _lib.stream_ptr(dev)
        lib = _lib.lib()
        cv = model.conv
        save = any(ctx.needs_input_grad)
        rec = {}
        with torch.cuda.device(dev):
            wprep = prepare_weights(model, st, dev, save)
            wfwd, wsplit = wprep["fwd"], wprep["w"]
            # conv.0 + ReLU + MaxPool2 -> NHWC [N, H/2, W/2, 32]
            H1, W1 = H // 2, W // 2
            a0 = _empty((N, H1, W1, 32), dev)
            # training: keep the arg-max positions / ReLU gates (4 bytes per pooled pixel and 8 channels) for conv.0's backward
            code0 = torch.empty((N * H1 * W1 * 4,), dtype=torch.int32, device=dev) if save else None
            call("ocrs_rec_conv0_fwd", ptr(x), N, H, W, ptr(cv["0"].weight), ptr(cv["0"].bias), ptr(a0), ptr(code0), st)

            def conv_bn_pool(inp, Hh, Ww, cin, conv, bn, ph, pw, mode, relu, kh=3, pad=1, out_strides=None, out=None):
                cout = conv.out_channels
                if _implicit_ok(cin, kh):
                    col, Ho, Wo = None, Hh, Ww
                    M = N * Ho * Wo
                    rows = lib.ocrs_gemm_stat_rows(M)
                    stats = _empty((rows, 2, cout), dev) if training else None
                    y = conv3x3(inp, N, Hh, Ww, cin, wfwd[id(conv.weight)], cout, st, stats=stats)
                else:
                    col, Ho, Wo = im2col(inp, N, Hh, Ww, cin, kh, kh, pad, pad, st)
                    M = N * Ho * Wo
                    rows = lib.ocrs_gemm_stat_rows(M)
                    stats = _empty((rows, 2, cout), dev) if training else None
                    y = gemm(col, col.shape[1], True, wfwd[id(conv.weight)], col.shape[1], True, M, cout, col.shape[1],
                             st, stats=stats, exact=EXACT_FWD, b_weight=True)
                bs = _bn_finalize(bn, stats, rows, M, training, relu, st, dev)
                Hp, Wp = Ho // ph, Wo // pw
                if out is None:
                    out = _empty((N, Hp, Wp, cout), dev)
                    out_strides = (Hp * Wp * cout, Wp * cout, cout)
                call("ocrs_rec_bn_act_pool_fwd", ptr(y), N, Ho, Wo, cout, ph, pw, mode, int(relu), ptr(bs.scale),
                     ptr(bs.shift), ptr(out), *out_strides, st)
                return out, Hp, Wp, dict(col=col, inp=inp, y=y, bs=bs, geom=(Ho, Wo, cout, ph, pw, mode, int(relu)),
                                         ostr=out_strides, inp_geom=(Hh, Ww, cin), k=(kh, pad))

            def conv_bias_relu(inp, Hh, Ww, cin, conv):
                if _implicit_ok(cin, 3):
                    a = conv3x3(inp, N, Hh, Ww, cin, wfwd[id(conv.weight)], conv.out_channels, st, bias=conv.bias, relu=True)
                    return a, dict(col=None, inp=inp, a=a, inp_geom=(Hh, Ww, cin))
                col, Ho, Wo = im2col(inp, N, Hh, Ww, cin, 3, 3, 1, 1, st)
                M = N * Ho * Wo
                a = gemm(col, col.shape[1], True, wfwd[id(conv.weight)], col.shape[1], True, M, conv.out_channels,
                         col.shape[1], st, bias=conv.bias, relu=True, exact=EXACT_FWD, b_weight=True)
                return a, dict(col=col, inp=inp, a=a, inp_geom=(Hh, Ww, cin))

            a3, H3, W3, rec["3"] = conv_bn_pool(a0, H1, W1, 32, cv["3"], cv["4"], 2, 2, 0, True)
            a7, rec["7"] = conv_bias_relu(a3, H3, W3, 64, cv["7"])
            a9, H9, W9, rec["9"] = conv_bn_pool(a7, H3, W3, 128, cv["9"], cv["10"], 2, 1, 0, True)
            a13, rec["13"] = conv_bias_relu(a9, H9, W9, 128, cv["13"])
            a15, H15, W15, rec["15"] = conv_bn_pool(a13, H9, W9, 128, cv["15"], cv["16"], 2, 1, 0, True)
            # conv.19 (2x2, pad 1) + BN + AvgPool((4,1)) stored directly as (T, N, C*H') with H' = 1
            Ho19, T = H15 + 1, W15 + 1
            if Ho19 // 4 != 1:
                raise RuntimeError(f"RecognitionModel expects input height 64 (got {H}): C*H' must stay 128")
            seq = _empty((T, N, 128), dev)
            _, _, _, rec["19"] = conv_bn_pool(a15, H15, W15, 128, cv["19"], cv["20"], 4, 1, 1, False, kh=2, pad=1,
                                              out_strides=(128, 0, N * 128), out=seq)
            # 2-layer bidirectional GRU
            gru = model.gru
            TN = T * N
            layer_in = seq
            gru_rec = []
            for layer in range(2):
                isz = 128 if layer == 0 else 512
                gi = []
                for sfx in ("", "_reverse"):
                    w_ih = getattr(gru, f"weight_ih_l{layer}{sfx}")
                    b_ih = getattr(gru, f"bias_ih_l{layer}{sfx}")
                    gi.append(gemm(layer_in, isz, True, wsplit.get(id(w_ih), w_ih), isz, True, TN, 768, isz, st, bias=b_ih, b_weight=True))
                out = _empty((T, N, 512), dev)
                gates = _empty((T, N, 2, 4, 256), dev)
                call("ocrs_gru_layer_fwd_persist", ptr(gi[0]), ptr(gi[1]), ptr(getattr(gru, f"weight_hh_l{layer}")),
                     ptr(getattr(gru, f"weight_hh_l{layer}_reverse")), ptr(getattr(gru, f"bias_hh_l{layer}")),
                     ptr(getattr(gru, f"bias_hh_l{layer}_reverse")), ptr(out), ptr(gates), T, N, st)
                gru_rec.append(dict(x=layer_in, out=out, gates=gates, isz=isz))
                layer_in = out
            lin = model.output[0]
            C = lin.out_features
            logits = gemm(layer_in, 512, True, wsplit.get(id(lin.weight), lin.weight), 512, True, TN, C, 512, st, bias=lin.bias, b_weight=True)
            lp = _empty((T, N, C), dev)
            call("ocrs_log_softmax_fwd", ptr(logits), ptr(lp), TN, C, st)
        if save:
            ctx.model = model
            ctx.rec = rec
            ctx.gru_rec = gru_rec
            ctx.misc = (x, code0, lp, N, H, W, T, C, bool(training))
            ctx.wprep = wprep
        return lp

    @staticmethod
    def backward(ctx, g_lp):
        model = ctx.model
        rec, gru_rec = ctx.rec, ctx.gru_rec
        x, code0, lp, N, H, W, T, C, training = ctx.misc
        wdg, wsplit, whhT_of = ctx.wprep["dg"], ctx.wprep["w"], ctx.wprep["whhT"]
        dev = lp.device
        st = _lib.stream_ptr(dev)
        lib = _lib.lib()
        cv, gru, lin = model.conv, model.gru, model.output[0]
        TN = T * N
        grads = {}
        g_lp = g_lp.contiguous().float()
        with torch.cuda.device(dev):
            ldd = (C + 3) // 4 * 4  # 16-byte row pitch (C = 97 -> 100): the two GEMMs below can then take dlog by TMA
            dlog = _empty((TN, ldd), dev)
            call("ocrs_log_softmax_bwd", ptr(lp), ptr(g_lp), ptr(dlog), TN, C, ldd, st)
            out1 = gru_rec[1]["out"]
            grads[id(lin.weight)] = gemm(dlog, ldd, False, out1, 512, False, C, 512, TN, st, split_ok=True, lazy={})
            grads[id(lin.bias)] = colsum(dlog, ldd, TN, C, st, dev)
            d_out = gemm(dlog, ldd, True, wsplit.get(id(lin.weight), lin.weight), 512, False, TN, 512, C, st, b_weight=True)
            for layer in (1, 0):
                r = gru_rec[layer]
                isz, xin, out, gates = r["isz"], r["x"], r["out"], r["gates"]
                names = [f"l{layer}", f"l{layer}_reverse"]
                whhT = [whhT_of[id(getattr(gru, "weight_hh_" + nm))] for nm in names]
                dgi = [_empty((TN, 768), dev) for _ in range(2)]
                dgh = [_empty((TN, 768), dev) for _ in range(2)]
                call("ocrs_gru_layer_bwd_persist", ptr(whhT[0]), ptr(whhT[1]), ptr(d_out), ptr(out), ptr(gates),
                     ptr(dgi[0]), ptr(dgi[1]), ptr(dgh[0]), ptr(dgh[1]), T, N, st)
                d_in = _empty((TN, isz), dev)
                for d, nm in enumerate(names):
                    w_ih = getattr(gru, "weight_ih_" + nm)
                    grads[id(w_ih)] = gemm(dgi[d], 768, False, xin, isz, False, 768, isz, TN, st, split_ok=True, lazy={})
                    grads[id(getattr(gru, "bias_ih_" + nm))] = colsum(dgi[d], 768, TN, 768, st, dev)
                    grads[id(getattr(gru, "bias_hh_" + nm))] = colsum(dgh[d], 768, TN, 768, st, dev)
                    # dW_hh = sum_t dgh[t]^T h_prev[t]; h_prev[t] = out[t-1] (fwd) / out[t+1] (reverse)
                    rows = (T - 1) * N
                    if rows > 0:
                        if d == 0:
                            pa = dgh[d].data_ptr() + 4 * N * 768
                            pb = out.data_ptr()
                        else:
                            pa = dgh[d].data_ptr()
                            pb = out.data_ptr() + 4 * (N * 512 + 256)
                        dwhh = gemm(pa, 768, False, pb, 512, False, 768, 256, rows, st, out=_empty((768, 256), dev),
                                    split_ok=True, lazy={})
                    else:
                        dwhh = torch.zeros((768, 256), device=dev)
                    grads[id(getattr(gru, "weight_hh_" + nm))] = dwhh
                    gemm(dgi[d], 768, True, wsplit.get(id(w_ih), w_ih), isz, False, TN, isz, 768, st, out=d_in, accumulate=(d == 1), b_weight=True)
                d_out = d_in
            d_seq = d_out  # [T, N, 128]

            def bn_pool_bwd(r, conv, bn, dout, dstr):
                Ho, Wo, cout, ph, pw, mode, relu = r["geom"]
                bs, y = r["bs"], r["y"]
                blocks = lib.ocrs_rec_pool_bwd_blocks()
                part = _empty((blocks, 2, cout), dev)
                call("ocrs_rec_bn_act_pool_bwd_reduce", ptr(y), N, Ho, Wo, cout, ph, pw, mode, relu, ptr(bs.scale),
                     ptr(bs.shift), ptr(bs.mean), ptr(bs.invstd), ptr(dout), *dstr, ptr(part), st)
                coef = _empty((5, cout), dev)
                call("ocrs_bn_bwd_finalize", ptr(part), blocks, cout, float(N * Ho * Wo), ptr(bn.weight), ptr(bs.mean),
                     ptr(bs.invstd), ptr(coef[0]), ptr(coef[1]), ptr(coef[2]), ptr(coef[3]), ptr(coef[4]), int(training), st)
                grads[id(bn.weight)] = coef[0]
                grads[id(bn.bias)] = coef[1]
                dy = _empty((N, Ho, Wo, cout), dev)
                call("ocrs_rec_bn_act_pool_bwd_apply", ptr(y), N, Ho, Wo, cout, ph, pw, mode, relu, ptr(bs.scale),
                     ptr(bs.shift), ptr(coef[2]), ptr(coef[3]), ptr(coef[4]), ptr(dout), *dstr, ptr(dy), st)
                return dy, Ho, Wo

            def conv_bwd(r, conv, dy, Ho, Wo, need_dx=True):
                """dy: [N, Ho, Wo, Cout] gradient of the raw conv output. Returns d(input) NHWC."""
                Hh, Ww, cin = r["inp_geom"]
                kh, pad = r.get("k", (3, 1))
                col = r["col"]
                cout = conv.out_channels
                if col is None:  # implicit GEMM: the weight gradient gathers the activation by TMA as well
                    M = N * Hh * Ww
                    dwp = conv3x3_wgrad(dy, r["inp"], N, Hh, Ww, cin, cout, st)
                else:
                    M, K = col.shape
                    dwp = gemm(dy, cout, False, col, K, False, cout, K, M, st, split_ok=True, lazy=dict(conv=(cin, kh * kh)))
                grads[id(conv.weight)] = dwp  # [Cout][(ky,kx,ci)] rows; grads.deliver writes the parameter layout
                if conv.bias is not None:
                    grads[id(conv.bias)] = colsum(dy, cout, M, cout, st, dev)
                if not need_dx:
                    return None
                if _implicit_ok(cout, kh):
                    return conv3x3(dy, N, Ho, Wo, cout, wdg[id(conv.weight)], cin, st)
                dcol, Hi, Wi = im2col(dy, N, Ho, Wo, cout, kh, kh, kh - 1 - pad, kh - 1 - pad, st)
                assert (Hi, Wi) == (Hh, Ww)
                return gemm(dcol, dcol.shape[1], True, wdg[id(conv.weight)], dcol.shape[1], True, N * Hh * Ww, cin,
                            dcol.shape[1], st, b_weight=True)

            r = rec["19"]
            dy, Ho, Wo = bn_pool_bwd(r, cv["19"], cv["20"], d_seq, r["ostr"])
            d_a15 = conv_bwd(r, cv["19"], dy, Ho, Wo)

            def stage(key_bn, key_relu, d_act, bn_name):
                """[conv+bias+ReLU] -> [conv+BN+ReLU+pool] pair, backwards. d_act: grad of the pooled output."""
                rb = rec[key_bn]
                Ho_, Wo_, cout, ph, pw, _, _ = rb["geom"]
                Hp, Wp = Ho_ // ph, Wo_ // pw
                dy_, Ho2, Wo2 = bn_pool_bwd(rb, cv[key_bn], cv[bn_name], d_act, (Hp * Wp * cout, Wp * cout, cout))
                d_prev = conv_bwd(rb, cv[key_bn], dy_, Ho2, Wo2)
                if key_relu is None:
                    return d_prev
                rr = rec[key_relu]
                call("ocrs_relu_bwd", ptr(rr["a"]), ptr(d_prev), d_prev.numel(), st)
                Hh, Ww, _ = rb["inp_geom"]
                return conv_bwd(rr, cv[key_relu], d_prev, Hh, Ww)

            d_a9 = stage("15", "13", d_a15, "16")
            d_a3 = stage("9", "7", d_a9, "10")
            d_a0 = stage("3", None, d_a3, "4")
            blocks = lib.ocrs_rec_conv0_bwd_blocks()
            part = _empty((blocks, 32, 10), dev)
            call("ocrs_rec_conv0_bwd", ptr(x), N, H, W, ptr(cv["0"].weight), ptr(cv["0"].bias), ptr(d_a0), ptr(code0), ptr(part), st)
            grads[id(cv["0"].weight)] = Partial(part, blocks, 288, ld=320, inner=(9, 10))
            grads[id(cv["0"].bias)] = Partial(part, blocks, 32, ld=320, off=9, inner=(1, 10))
            param_grads = deliver(list(model.parameters()), grads, st)
        ctx.rec = ctx.gru_rec = ctx.misc = ctx.wprep = None
        return (None, None) + tuple(param_grads)


_env_mode_applied = False


def recognition_forward(model, x: torch.Tensor) -> torch.Tensor:
    global _env_mode_applied
    if not _env_mode_applied:
        _env_mode_applied = True
        if os.environ.get("OCRS_PRECISION", "parity") == "tf32":
            set_precision("tf32")
    if not x.is_cuda:
        raise RuntimeError("ocrs_models_b200.RecognitionModel has no CPU path: input must be a CUDA tensor")
    if x.dim() != 4 or x.shape[1] != 1:
        raise RuntimeError(f"expected (N, 1, 64, W) input, got {tuple(x.shape)}")
    if x.shape[3] < 4:
        raise RuntimeError("input width must be at least 4")
    if x.requires_grad and torch.is_grad_enabled():
        # conv.0's backward kernel produces weight/bias gradients only; refuse rather than return a silent zero
        raise RuntimeError("ocrs_models_b200.RecognitionModel does not compute the gradient w.r.t. the input image "
                           "(x.requires_grad must be False)")
    if model.__dict__.get("_ocrs_checked") != x.device:
        _lib.check_module_tensors(model, x.device, "RecognitionModel")
        model.__dict__["_ocrs_checked"] = x.device
    x = x.float().contiguous()
    return _RecFunction.apply(model, x, *model.parameters())
