"""Detection kernels through the C-ABI vs the oracle (oracle/functional.py, CPU, fp64 = ground truth)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import functional as O

pytestmark = pytest.mark.gpu


def _stream():
    from ocrs_models_b200 import _lib

    return _lib.stream_ptr(torch.device("cuda:0"))


def _rand_xf(C, g):
    sc = torch.randn(C, generator=g) * 0.5 + 1.0
    sc[::3] *= -1  # negative BN scales occur in trained models
    sh = torch.randn(C, generator=g) * 0.3
    lo = torch.zeros(C)
    return sc, sh, lo


def _apply_xf(x, xf):
    sc, sh, lo = xf
    return torch.maximum(x * sc[None, :, None, None] + sh[None, :, None, None], lo[None, :, None, None])


@pytest.mark.parametrize("N,cin,cout,H,W,with_xf", [(2, 1, 8, 37, 45, False), (2, 8, 16, 64, 64, True), (1, 16, 16, 33, 70, True), (2, 64, 32, 9, 12, True), (1, 128, 256, 5, 3, True), (1, 24, 40, 17, 19, True),
                                                         # TMA-addressable shapes (W % 4 == 0): csrc/det_tma.cu
                                                         (2, 1, 8, 40, 36, False), (1, 16, 16, 68, 100, True), (2, 32, 16, 96, 64, True),
                                                         (1, 64, 32, 32, 40, True), (1, 16, 32, 36, 44, True), (1, 256, 128, 8, 16, True),
                                                         (3, 8, 8, 33, 132, True),
                                                         # >= 64 channels: batched tcgen05 GEMMs (csrc/gemm_tc.cu)
                                                         (2, 64, 64, 16, 16, True), (3, 32, 64, 24, 20, True), (2, 128, 256, 8, 8, True)])
def test_separable_block_fwd_bwd(N, cin, cout, H, W, with_xf):
    from ocrs_models_b200.det_engine import View, _Sep, new_view
    from ocrs_models_b200.models import _separable

    g = torch.Generator().manual_seed(cin * 100 + cout)
    torch.manual_seed(cin * 100 + cout)
    mod = _separable(cin, cout)
    with torch.no_grad():
        mod.seq[2].weight.copy_(torch.randn(cout, generator=g) * 0.5 + 1)
        mod.seq[2].bias.copy_(torch.randn(cout, generator=g) * 0.2)
    x = torch.randn(N, cin, H, W, generator=g)
    xf = _rand_xf(cin, g) if with_xf else None
    d_a = torch.randn(N, cout, H, W, generator=g)

    # oracle in fp64
    sd = {"b." + k: v.detach().double().clone() for k, v in mod.state_dict().items()}
    for k in ("b.seq.0.weight", "b.seq.1.weight", "b.seq.2.weight", "b.seq.2.bias"):
        sd[k].requires_grad_(True)
    xa = (_apply_xf(x, xf) if xf else x).double().requires_grad_(True)
    nb = {}
    out = O._depthwise_block(sd, "b", xa, True, nb)
    out.backward(d_a.double())

    modc = _separable(cin, cout).cuda()
    modc.load_state_dict(mod.state_dict())
    blk = _Sep(modc)
    xd = x.cuda()
    xfd = tuple(a.cuda() for a in xf) if xf else None
    inp = View(xd, 0, cin * H * W, cin, H, W, xfd)
    recs = {}
    y = blk.forward(inp, N, True, _stream(), save=recs)
    act = _apply_xf(y.t, y.xf)
    assert rel_l2(act, out) < 1e-5
    assert rel_l2(modc.seq[2].running_mean, nb["b.seq.2.running_mean"]) < 1e-5
    assert rel_l2(modc.seq[2].running_var, nb["b.seq.2.running_var"]) < 1e-5
    assert int(modc.seq[2].num_batches_tracked) == 1
    dad = d_a.cuda()
    grads, dx = blk.backward(recs, View(dad, 0, cout * H * W, cout, H, W), N, _stream(), None)
    torch.cuda.synchronize()
    assert rel_l2(dx.t, xa.grad) < 2e-4, "dx"
    names = ["b.seq.0.weight", "b.seq.1.weight", "b.seq.2.weight", "b.seq.2.bias"]
    from ocrs_models_b200.grads import materialize

    for gname, got in zip(names, grads):
        want = sd[gname].grad
        got = materialize(got, tuple(want.shape), _stream())  # weight gradients leave the block as partial rows
        err = (got.cpu().double() - want).norm() / max(want.norm(), 1e-3 * d_a.numel() ** 0.5)
        assert err < 1e-3, (gname, float(err))


def test_dw_bwd_accumulate_into_strided_view():
    from ocrs_models_b200._lib import call, ptr

    g = torch.Generator().manual_seed(1)
    N, C, H, W = 2, 8, 21, 34
    gr = torch.randn(N, C, H, W, generator=g).cuda()
    w = torch.randn(C, 1, 3, 3, generator=g).cuda()
    base = torch.randn(N, 2 * C, H, W, generator=g).cuda()
    ref = base.clone()
    ref[:, C:] += F.conv_transpose2d(gr, w, padding=1, groups=C)
    dst_off = C * H * W
    call("ocrs_det_dw_bwd", ptr(gr), C * H * W, ptr(gr), C * H * W, N, C, H, W, None, None, None, ptr(w),
         base.data_ptr() + 4 * dst_off, 2 * C * H * W, 1, None, _stream())
    assert rel_l2(base, ref) < 1e-6


@pytest.mark.parametrize("N,cin,cout,H,W", [(2, 8, 16, 40, 36), (1, 1, 8, 17, 64), (2, 32, 16, 20, 72), (1, 16, 40, 9, 132), (1, 24, 8, 33, 32)])
def test_tma_pw_wgrad_matches_fp64(N, cin, cout, H, W):
    """ocrs_det_sep_pw_wgrad (TMA-staged tiles, mma.sync 3xTF32) vs an fp64 einsum of dy and the depthwise output."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    lib = _lib.lib()
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(N, cin, H, W, generator=g)
    y = torch.randn(N, cout, H, W, generator=g)
    d_a = torch.randn(N, cout, H, W, generator=g)
    wdw = torch.randn(cin, 1, 3, 3, generator=g)
    xf, yxf = _rand_xf(cin, g), _rand_xf(cout, g)
    k1, k2, k3 = (torch.randn(cout, generator=g) for _ in range(3))
    xa = _apply_xf(x, xf).double()
    dwout = F.conv2d(xa, wdw.double(), padding=1, groups=cin)
    act = (y.double() * yxf[0][None, :, None, None] + yxf[1][None, :, None, None]) > 0
    dy = k1[None, :, None, None] * (d_a.double() * act) + k2[None, :, None, None] * y.double() + k3[None, :, None, None]
    ref = torch.einsum("nohw,nihw->oi", dy, dwout)
    workers = lib.ocrs_det_sep_pw_wgrad_workers(N, H, W, cout, cin)
    part = torch.full((workers, cout, cin), float("nan"), device="cuda")
    dev = [t.cuda() for t in (d_a, y, *yxf, k1, k2, k3, x, *xf, wdw)]
    call("ocrs_det_sep_pw_wgrad", ptr(dev[0]), cout * H * W, ptr(dev[1]), cout * H * W, N, cout, H, W, ptr(dev[2]), ptr(dev[3]),
         ptr(dev[4]), ptr(dev[5]), ptr(dev[6]), ptr(dev[7]), ptr(dev[8]), cin * H * W, cin, ptr(dev[9]), ptr(dev[10]), ptr(dev[11]),
         ptr(dev[12]), ptr(part), _stream())
    torch.cuda.synchronize()
    assert rel_l2(part.double().sum(0), ref) < 2e-5


@pytest.mark.parametrize("N,cin,cout,HW", [(2, 8, 16, 40 * 36), (3, 1, 8, 36 * 45), (1, 32, 16, 1000), (2, 16, 40, 132), (1, 24, 8, 8), (2, 16, 16, 128 * 9)])
def test_pw_wgrad_from_saved_depthwise_output(N, cin, cout, HW):
    """ocrs_det_pw_wgrad_saved: both operands streamed in mma fragment layout; any pixel count (tails masked)."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    lib = _lib.lib()
    g = torch.Generator().manual_seed(cin * 11 + cout)
    dwo = torch.randn(N, cin, HW, generator=g)
    y = torch.randn(N, cout, HW, generator=g)
    d_a = torch.randn(N, cout, HW, generator=g)
    yxf = _rand_xf(cout, g)
    k1, k2, k3 = (torch.randn(cout, generator=g) for _ in range(3))
    act = (y.double() * yxf[0][None, :, None] + yxf[1][None, :, None]) > 0
    dy = k1[None, :, None] * (d_a.double() * act) + k2[None, :, None] * y.double() + k3[None, :, None]
    ref = torch.einsum("nop,nip->oi", dy, dwo.double())
    workers = lib.ocrs_det_pw_wgrad_saved_workers(N, HW, cout, cin)
    part = torch.full((workers, cout, cin), float("nan"), device="cuda")
    dev = [t.cuda() for t in (d_a, y, *yxf, k1, k2, k3, dwo)]
    wpw = torch.randn(cout, cin, generator=g)
    fuse = cout <= 16
    gout = torch.full((N, cin, HW), float("nan"), device="cuda")
    wd = wpw.cuda()
    call("ocrs_det_pw_wgrad_saved", ptr(dev[0]), cout * HW, ptr(dev[1]), cout * HW, N, cout, HW, ptr(dev[2]), ptr(dev[3]), ptr(dev[4]),
         ptr(dev[5]), ptr(dev[6]), ptr(dev[7]), ptr(dev[8]), cin, ptr(part), ptr(wd) if fuse else None, ptr(gout) if fuse else None,
         cin * HW, _stream())
    torch.cuda.synchronize()
    assert rel_l2(part.double().sum(0), ref) < 2e-5
    if fuse:  # fused 1x1 data gradient
        assert rel_l2(gout, torch.einsum("oi,nop->nip", wpw.double(), dy)) < 2e-6


@pytest.mark.parametrize("N,cin,HW", [(2, 32, 64 * 48), (1, 16, 1000), (3, 64, 132), (1, 32, 64), (2, 32, 8)])
def test_pw_wgrad_saved_32_output_channels_with_data_gradient(N, cin, HW):
    """ocrs_det_pw_wgrad_saved32: the blocks with 32 output channels; weight gradient partials and the fused 1x1 data
    gradient from one staged tile (what ocrs_det_pw_wgrad_saved + ocrs_det_pwT_bwd compute in two passes)."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    lib = _lib.lib()
    cout = 32
    g = torch.Generator().manual_seed(cin * 7 + HW)
    dwo = torch.randn(N, cin, HW, generator=g)
    y = torch.randn(N, cout, HW, generator=g)
    d_a = torch.randn(N, cout, HW, generator=g)
    yxf = _rand_xf(cout, g)
    k1, k2, k3 = (torch.randn(cout, generator=g) for _ in range(3))
    wpw = torch.randn(cout, cin, generator=g)
    act = (y.double() * yxf[0][None, :, None] + yxf[1][None, :, None]) > 0
    dy = k1[None, :, None] * (d_a.double() * act) + k2[None, :, None] * y.double() + k3[None, :, None]
    workers = lib.ocrs_det_pw_wgrad_saved32_workers(N, HW, cin)
    part = torch.full((workers, cout, cin), float("nan"), device="cuda")
    # d_a and y as channel slices of wider buffers (sample stride > 32 planes), g into a slice too
    da_buf = torch.zeros(N, cout + 4, HW, device="cuda")
    da_buf[:, :cout] = d_a.cuda()
    y_buf = torch.zeros(N, cout + 8, HW, device="cuda")
    y_buf[:, :cout] = y.cuda()
    g_buf = torch.full((N, cin + 4, HW), float("nan"), device="cuda")
    dev = [t.cuda() for t in (*yxf, k1, k2, k3, dwo, wpw)]
    call("ocrs_det_pw_wgrad_saved32", ptr(da_buf), (cout + 4) * HW, ptr(y_buf), (cout + 8) * HW, N, HW, ptr(dev[0]), ptr(dev[1]),
         ptr(dev[2]), ptr(dev[3]), ptr(dev[4]), ptr(dev[5]), ptr(dev[6]), cin, ptr(part), ptr(dev[7]), ptr(g_buf), (cin + 4) * HW,
         _stream())
    torch.cuda.synchronize()
    assert rel_l2(part.double().sum(0), torch.einsum("nop,nip->oi", dy, dwo.double())) < 2e-5
    assert rel_l2(g_buf[:, :cin], torch.einsum("oi,nop->nip", wpw.double(), dy)) < 2e-6
    assert torch.isnan(g_buf[:, cin:]).all()


@pytest.mark.parametrize("N,C,H,W,acc", [(2, 8, 40, 36, True), (1, 5, 70, 132, False), (2, 16, 32, 64, True), (1, 1, 9, 8, False)])
def test_tma_dw_bwd_with_fused_upstream_bn_reduction(N, C, H, W, acc):
    """ocrs_det_sep_dw_bwd: dx (accumulated into a strided channel slice), dw weight gradient, and the BatchNorm-backward
    sums of the block that produced x, vs the separate kernels (ocrs_det_dw_bwd + ocrs_bnrelu_bwd_reduce) and fp64."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    lib = _lib.lib()
    g = torch.Generator().manual_seed(C * H)
    gr = torch.randn(N, C, H, W, generator=g)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(C, 1, 3, 3, generator=g)
    xf = _rand_xf(C, g)
    mean, invstd = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    base = torch.randn(N, 2 * C, H, W, generator=g)
    xa = _apply_xf(x, xf).double()
    wd = w.double().requires_grad_(True)
    xa.requires_grad_(True)
    out = F.conv2d(xa, wd, padding=1, groups=C)
    out.backward(gr.double())
    ref_dx = xa.grad + (base[:, C:].double() if acc else 0)
    dz = ref_dx * (xa > 0)
    yhat = (x.double() - mean[None, :, None, None]) * invstd[None, :, None, None]
    ref_bn = torch.stack([dz.sum((0, 2, 3)), (dz * yhat).sum((0, 2, 3))])
    grd, xd, wdv, based = gr.cuda(), x.cuda(), w.cuda(), base.cuda()
    xfd = [t.cuda() for t in xf]
    rows = lib.ocrs_det_sep_dw_bwd_rows(N, H, W, C)
    wpart = torch.full((rows, C, 9), float("nan"), device="cuda")
    bnpart = torch.full((rows, 2, C), float("nan"), device="cuda")
    md, isd = mean.cuda(), invstd.cuda()
    call("ocrs_det_sep_dw_bwd", ptr(grd), C * H * W, ptr(xd), C * H * W, N, C, H, W, *[ptr(t) for t in xfd], ptr(wdv),
         based.data_ptr() + 4 * C * H * W, 2 * C * H * W, int(acc), ptr(wpart), ptr(md), ptr(isd), ptr(bnpart), _stream())
    torch.cuda.synchronize()
    assert rel_l2(based[:, C:], ref_dx) < 1e-6
    assert torch.equal(based[:, :C].cpu(), base[:, :C])  # the other half of the concat gradient is untouched
    assert rel_l2(wpart.double().sum(0).reshape(C, 1, 3, 3), wd.grad) < 1e-5
    assert rel_l2(bnpart.double().sum(0), ref_bn) < 1e-5


@pytest.mark.parametrize("H,W", [(64, 64), (37, 51), (2, 3)])
def test_pool_fwd_bwd(H, W):
    from ocrs_models_b200._lib import call, ptr

    g = torch.Generator().manual_seed(H)
    N, C = 2, 5
    x = torch.randn(N, C, H, W, generator=g)
    xf = _rand_xf(C, g)
    a = _apply_xf(x, xf).requires_grad_(True)
    ref = F.max_pool2d(a, 2)
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout)
    xd, xfd = x.cuda(), [t.cuda() for t in xf]
    Ho, Wo = H // 2, W // 2
    out = torch.empty(N, C, Ho, Wo, device="cuda")
    call("ocrs_det_pool2_fwd", ptr(xd), C * H * W, N, C, H, W, *[ptr(t) for t in xfd], ptr(out), C * Ho * Wo, _stream())
    assert rel_l2(out, ref) < 1e-6  # fmaf vs mul+add in the test's own transform
    din = torch.full((N, C, H, W), 7.0, device="cuda")
    doutd = dout.cuda()
    call("ocrs_det_pool2_bwd", ptr(xd), C * H * W, N, C, H, W, *[ptr(t) for t in xfd], ptr(doutd), C * Ho * Wo,
         ptr(din), C * H * W, _stream())
    assert rel_l2(din, a.grad) < 1e-6


@pytest.mark.parametrize("N,cin,cout,h,w,Hs,Ws", [(2, 16, 8, 20, 24, 40, 48), (1, 256, 128, 3, 2, 7, 5), (2, 32, 16, 9, 7, 18, 15), (1, 24, 12, 5, 6, 11, 13),
                                                  (2, 16, 8, 9, 68, 19, 136), (1, 32, 32, 6, 16, 12, 32), (1, 8, 20, 5, 80, 11, 160)])
def test_convt_fwd_bwd(N, cin, cout, h, w, Hs, Ws):
    from ocrs_models_b200._lib import call, lib, ptr

    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(N, cin, h, w, generator=g)
    xf = _rand_xf(cin, g)
    wt = (torch.randn(cin, cout, 3, 3, generator=g) * 0.1).double().requires_grad_(True)
    b = torch.randn(cout, generator=g).double().requires_grad_(True)
    a = _apply_xf(x, xf).double().requires_grad_(True)
    ref = F.conv_transpose2d(a, wt, b, stride=2)[:, :, :Hs, :Ws]
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    xd, xfd = x.cuda(), [ptr(t.cuda()) for t in xf]
    keep = [t.cuda() for t in xf]
    xfd = [ptr(t) for t in keep]
    wd, bd = wt.detach().float().cuda(), b.detach().float().cuda()
    # write into the lower half of a wider concat buffer
    cat = torch.zeros(N, cout + 3, Hs, Ws, device="cuda")
    call("ocrs_det_convt_fwd", ptr(xd), cin * h * w, N, cin, h, w, *xfd, ptr(wd), ptr(bd), cout, ptr(cat),
         (cout + 3) * Hs * Ws, Hs, Ws, _stream())
    assert rel_l2(cat[:, :cout], ref) < 1e-5
    assert cat[:, cout:].abs().max().item() == 0
    dcat = torch.zeros_like(cat)
    dcat[:, :cout] = dout.cuda()
    dx = torch.empty(N, cin, h, w, device="cuda")
    call("ocrs_det_convt_bwd_data", ptr(dcat), (cout + 3) * Hs * Ws, N, cout, Hs, Ws, ptr(wd), cin, h, w, ptr(dx),
         cin * h * w, _stream())
    assert rel_l2(dx, a.grad) < 1e-5
    workers = lib().ocrs_det_convt_wgrad_workers(N, h, w)
    part = torch.empty(workers, cin, cout, 3, 3, device="cuda")
    call("ocrs_det_convt_wgrad", ptr(xd), cin * h * w, N, cin, h, w, *xfd, ptr(dcat), (cout + 3) * Hs * Ws, cout, Hs,
         Ws, ptr(part), _stream())
    dw = torch.empty(cin, cout, 3, 3, device="cuda")
    call("ocrs_finalize_partials", ptr(part), workers, dw.numel(), ptr(dw), _stream())
    assert rel_l2(dw, wt.grad) < 1e-5
    cs = (cout + 3) * Hs * Ws
    if lib().ocrs_det_convt_wgrad_staged_ok(ptr(xd), cin * h * w, h, w, ptr(dcat), cs, Hs, Ws):
        # cp.async-staged version (csrc/det_tma.cu) on TMA-addressable shapes
        workers = lib().ocrs_det_convt_wgrad_staged_workers(N, h, w, cin, cout)
        part = torch.full((workers, cin, cout, 3, 3), float("nan"), device="cuda")
        call("ocrs_det_convt_wgrad_staged", ptr(xd), cin * h * w, N, cin, h, w, *xfd, ptr(dcat), cs, cout, Hs, Ws, ptr(part), _stream())
        assert rel_l2(part.double().sum(0), wt.grad) < 1e-5, "staged wgrad"
    rows = lib().ocrs_reduce_rows(N, Hs * Ws)
    bp = torch.empty(rows, cout, device="cuda")
    call("ocrs_plane_sum", ptr(dcat), (cout + 3) * Hs * Ws, N, cout, Hs * Ws, ptr(bp), _stream())
    db = torch.empty(cout, device="cuda")
    call("ocrs_finalize_partials", ptr(bp), rows, cout, ptr(db), _stream())
    assert rel_l2(db, b.grad) < 1e-5


@pytest.mark.parametrize("N,cin,cout,h,w,Hs,Ws", [(2, 64, 32, 8, 8, 16, 16), (1, 256, 128, 8, 12, 17, 25), (3, 128, 64, 10, 10, 20, 21),
                                                  (2, 64, 32, 16, 4, 33, 8)])
def test_convt_as_tcgen05_gemms(N, cin, cout, h, w, Hs, Ws):
    """ConvTranspose2d with >= 64 input channels: activate -> batched GEMM -> col2im, im2col -> two batched GEMMs
    (det_engine._convt_forward_tc / _convt_backward_tc) vs fp64 conv_transpose2d (models.py:76-89, incl. the crop)."""
    from ocrs_models_b200 import det_engine as E

    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(N, cin, h, w, generator=g)
    xf = _rand_xf(cin, g)
    wt = (torch.randn(cin, cout, 3, 3, generator=g) * 0.1).double().requires_grad_(True)
    b = torch.randn(cout, generator=g).double().requires_grad_(True)
    a = _apply_xf(x, xf).double().requires_grad_(True)
    ref = F.conv_transpose2d(a, wt, b, stride=2)[:, :, :Hs, :Ws]
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    t = torch.nn.ConvTranspose2d(cin, cout, 3, stride=2).cuda()
    with torch.no_grad():
        t.weight.copy_(wt.detach().float())
        t.bias.copy_(b.detach().float())
    up = E.View(x.cuda(), 0, cin * h * w, cin, h, w, tuple(v.cuda() for v in xf))
    assert E._convt_tc_ok(t, up, cout)
    cat = torch.zeros(N, cout + 3, Hs, Ws, device="cuda")
    out = E.View(cat, 0, (cout + 3) * Hs * Ws, cout, Hs, Ws)
    xa = E._convt_forward_tc(t, up, N, cout, out, True, _stream())
    assert rel_l2(xa.view(N, cin, h, w), a.detach()) < 1e-6
    assert rel_l2(cat[:, :cout], ref) < 2e-6
    assert cat[:, cout:].abs().max().item() == 0
    dcat = torch.zeros_like(cat)
    dcat[:, :cout] = dout.cuda()
    dlo = E.View(dcat, 0, (cout + 3) * Hs * Ws, cout, Hs, Ws)
    d_up, dw = E._convt_backward_tc(t, xa, dlo, N, cin, h, w, _stream())
    assert rel_l2(d_up.t, a.grad) < 2e-6
    from ocrs_models_b200.grads import materialize

    assert rel_l2(materialize(dw, (cin, cout, 3, 3), _stream()), wt.grad) < 2e-6


def _loss_case(p, t):
    from ocrs_models_b200 import balanced_cross_entropy_loss

    pd = p.cuda().requires_grad_(True)
    lo = balanced_cross_entropy_loss(pd, t.cuda())
    lo.backward()
    pc = p.double().requires_grad_(True)
    lr = O.balanced_cross_entropy_loss(pc, t.double())
    lr.backward()
    return lo.detach().cpu(), pd.grad.cpu(), lr.detach(), pc.grad


@pytest.mark.parametrize("shape,frac", [((2, 1, 64, 80), 0.1), ((1, 1, 301, 257), 0.7), ((3, 1, 33, 33), 0.5)])
def test_balanced_bce_vs_oracle(shape, frac):
    g = torch.Generator().manual_seed(int(frac * 100))
    p = torch.sigmoid(torch.randn(shape, generator=g) * 3)
    t = (torch.rand(shape, generator=g) < frac).float()
    t.view(-1)[5] = 1.2   # out-of-range targets are clamped (train_detection.py:249)
    t.view(-1)[9] = -0.1
    t.view(-1)[11] = 0.5  # in neither class
    lo, go, lr, gr = _loss_case(p, t)
    assert abs(lo - lr) < 1e-5 * abs(lr)
    assert rel_l2(go, gr) < 1e-5


def test_balanced_bce_saturated_and_ties():
    g = torch.Generator().manual_seed(4)
    shape = (1, 1, 40, 40)
    p = torch.sigmoid(torch.randn(shape, generator=g))
    t = (torch.rand(shape, generator=g) < 0.2).float()
    p.view(-1)[:7] = 0.0  # log clamp at -100
    p.view(-1)[7:12] = 1.0
    t.view(-1)[:12] = torch.tensor([1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1, 0]).float()
    lo, go, lr, gr = _loss_case(p, t)
    assert abs(lo - lr) < 1e-5 * abs(lr)
    # tie at the threshold: quantised losses -> many equal values; loss must still agree
    p2 = (torch.randint(1, 8, shape, generator=g).float() / 8.0)
    lo, go, lr, gr = _loss_case(p2, t)
    assert abs(lo - lr) < 1e-5 * abs(lr)
    assert abs(go.sum() - gr.sum()) < 1e-4 * gr.abs().sum()


def test_balanced_bce_nan_prediction_gives_nan_loss():
    from ocrs_models_b200 import balanced_cross_entropy_loss

    g = torch.Generator().manual_seed(0)
    p = torch.rand(1, 1, 16, 16, generator=g) * 0.8 + 0.1
    p[0, 0, 3, 3] = float("nan")
    t = (torch.rand(1, 1, 16, 16, generator=g) < 0.3).float()
    assert torch.isnan(balanced_cross_entropy_loss(p.cuda(), t.cuda()))  # a diverged run must not look healthy


def test_balanced_bce_degenerate_no_positives():
    from ocrs_models_b200 import balanced_cross_entropy_loss

    p = torch.full((1, 1, 8, 8), 0.3, device="cuda", requires_grad=True)
    lo = balanced_cross_entropy_loss(p, torch.zeros(1, 1, 8, 8, device="cuda"))
    lo.backward()
    assert torch.isnan(lo)  # mean of an empty selection, as in the reference
    assert p.grad.abs().max().item() == 0


def _det_models(seed=1234):
    from ocrs_models_b200 import DetectionModel

    torch.manual_seed(seed)
    m = DetectionModel()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # move BN affine away from (1, 0) so its gradients are exercised
        for k, v in m.named_parameters():
            if ".seq.2." in k:
                v.add_(torch.randn(v.shape, generator=g) * 0.2)
    return m


@pytest.mark.parametrize("N,H,W", [(2, 96, 80), (1, 200, 152), (2, 64, 64), (1, 256, 320)])
def test_full_model_train_step_vs_oracle(N, H, W):
    from ocrs_models_b200 import balanced_cross_entropy_loss

    m = _det_models()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 1, H, W, generator=g) - 0.5
    mask = (torch.rand(N, 1, H, W, generator=g) < 0.1).float()
    batch = {"image": x, "mask": mask}
    out64, loss64, g64, nb64 = O.train_step_grads("det", sd, batch, torch.float64)
    out32, loss32, g32, _ = O.train_step_grads("det", sd, batch, torch.float32)

    m = m.cuda().train()
    y = m(x.cuda())
    loss = balanced_cross_entropy_loss(y, mask.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(y, out64) < 1e-4, "probabilities"
    assert abs(loss.item() - loss64.item()) < 1e-4 * abs(loss64.item())
    ours = {k: p.grad for k, p in m.named_parameters()}
    gn = torch.sqrt(sum((v.double() ** 2).sum() for v in g64.values()))
    err = torch.sqrt(sum(((ours[k].cpu().double() - g64[k]) ** 2).sum() for k in g64)) / gn
    err32 = torch.sqrt(sum(((g32[k].double() - g64[k]) ** 2).sum() for k in g64)) / gn
    print(f"global grad rel-L2: ours {err:.3e}  fp32-oracle {err32:.3e}")
    assert err < 1e-3
    floor = gn / len(g64) ** 0.5
    for k in g64:
        e = (ours[k].cpu().double() - g64[k]).norm()
        e32 = (g32[k].double() - g64[k]).norm()
        assert e <= max(1e-3 * g64[k].norm(), 1e-3 * floor, 10 * e32), (k, float(e), float(g64[k].norm()), float(e32))
    for k, v in nb64.items():
        if v.is_floating_point():
            assert rel_l2(m.state_dict()[k], v) < 1e-4, k
        else:
            assert int(m.state_dict()[k]) == int(v)


def test_full_model_golden_and_eval_mode(golden):
    from ocrs_models_b200 import DetectionModel, balanced_cross_entropy_loss

    gold = golden("det_96x80")
    torch.manual_seed(1234)
    m = DetectionModel().cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 96, 80, generator=g) - 0.5
    mask = (torch.rand(2, 1, 96, 80, generator=g) < 0.1).float()
    y = m(x.cuda())
    loss = balanced_cross_entropy_loss(y, mask.cuda())
    loss.backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), gold["y"], rtol=1e-3, atol=1e-5)
    assert abs(loss.item() - gold["loss"]) < 1e-4 * gold["loss"]
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())).item()
    assert abs(gn - gold["grad_norm"]) < 1e-3 * gold["grad_norm"]
    # eval mode uses the running statistics and records nothing for backward
    m.eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.inference_mode():
        ye = m(x.cuda())
    ref = O.det_forward(sd, x, training=False)
    assert rel_l2(ye, ref) < 1e-4


def test_rejects_cpu_input():
    from ocrs_models_b200 import DetectionModel

    with pytest.raises(RuntimeError, match="no CPU path"):
        DetectionModel()(torch.zeros(1, 1, 64, 64))
