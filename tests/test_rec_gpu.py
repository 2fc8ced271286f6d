"""Recognition kernels through the C-ABI vs the oracle (oracle/functional.py, CPU, fp64 = ground truth)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import functional as O

pytestmark = pytest.mark.gpu


def _st():
    from ocrs_models_b200 import _lib

    return _lib.stream_ptr(torch.device("cuda:0"))


@pytest.mark.parametrize("M,N,K", [(300, 97, 512), (128, 128, 64), (1000, 64, 288), (37, 200, 20), (768, 256, 1930)])
@pytest.mark.parametrize("ak,bk", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("backend", ["tc", "simt"])
def test_gemm_layouts(M, N, K, ak, bk, backend, monkeypatch):
    from ocrs_models_b200 import rec_engine
    from ocrs_models_b200.rec_engine import gemm

    monkeypatch.setattr(rec_engine, "GEMM_BACKEND", backend)

    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ B.double().t()
    Ad = (A if ak else A.t().contiguous()).cuda()
    Bd = (B if bk else B.t().contiguous()).cuda()
    tol = 3e-6  # 4-accumulator 3xTF32 (tensor-core accumulation truncates) and fp32 FMA alike
    out = gemm(Ad, Ad.shape[1], ak, Bd, Bd.shape[1], bk, M, N, K, _st())
    assert rel_l2(out, ref) < tol
    out = gemm(Ad, Ad.shape[1], ak, Bd, Bd.shape[1], bk, M, N, K, _st(), bias=bias.cuda(), relu=True)
    assert rel_l2(out, torch.relu(ref + bias.double())) < tol
    out2 = out.clone()
    gemm(Ad, Ad.shape[1], ak, Bd, Bd.shape[1], bk, M, N, K, _st(), out=out2, accumulate=True)
    assert rel_l2(out2, out.cpu().double() + ref) < tol
    out = gemm(Ad, Ad.shape[1], ak, Bd, Bd.shape[1], bk, M, N, K, _st(), split_ok=True)
    assert rel_l2(out, ref) < tol


@pytest.mark.parametrize("M,N,K,bk", [(300, 97, 512, True), (1000, 64, 288, True), (640, 768, 128, True), (500, 128, 768, False)])
def test_gemm_presplit_weights_bit_exact(M, N, K, bk, monkeypatch):
    """B (a weight matrix) split once in HBM by ocrs_split_tf32 must give the very same bits as the in-kernel
    conversion: the split kernel and the converter warps use the same rounding."""
    from ocrs_models_b200 import rec_engine
    from ocrs_models_b200.rec_engine import Split, gemm

    monkeypatch.setattr(rec_engine, "PRESPLIT", False)  # plain tensors -> in-kernel conversion (the reference here)
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g)
    Bd = (B if bk else B.t().contiguous()).cuda()
    bias = torch.randn(N, generator=g).cuda()
    sp = Split(Bd, _st())
    assert torch.equal(sp.hi + sp.lo, Bd) and torch.equal(sp.hi.view(torch.int32) & 0x1FFF, torch.zeros_like(sp.hi, dtype=torch.int32))
    ref = gemm(A, K, True, Bd, Bd.shape[1], bk, M, N, K, _st(), bias=bias, relu=True)
    out = gemm(A, K, True, sp, Bd.shape[1], bk, M, N, K, _st(), bias=bias, relu=True)
    assert torch.equal(out, ref)
    assert rel_l2(out, torch.relu(A.double().cpu() @ B.double().t() + bias.double().cpu())) < 3e-6


def test_conv3x3_presplit_weights_bit_exact(monkeypatch):
    from ocrs_models_b200 import rec_engine
    from ocrs_models_b200.rec_engine import Split, _w_fwd, conv3x3

    monkeypatch.setattr(rec_engine, "PRESPLIT", False)
    g = torch.Generator().manual_seed(11)
    N, H, W, cin, cout = 2, 16, 200, 64, 128
    x = torch.randn(N, H, W, cin, generator=g).cuda()
    wp = _w_fwd((torch.randn(cout, cin, 3, 3, generator=g) * 0.1).cuda())
    ref = conv3x3(x, N, H, W, cin, wp, cout, _st())
    out = conv3x3(x, N, H, W, cin, Split(wp, _st()), cout, _st())
    assert torch.equal(out, ref)


@pytest.mark.parametrize("backend", ["tc", "simt"])
def test_gemm_column_stats(backend, monkeypatch):
    from ocrs_models_b200 import _lib, rec_engine
    from ocrs_models_b200.rec_engine import gemm

    monkeypatch.setattr(rec_engine, "GEMM_BACKEND", backend)

    g = torch.Generator().manual_seed(0)
    M, N, K = 700, 64, 96
    A, B = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    rows = _lib.lib().ocrs_gemm_stat_rows(M)
    stats = torch.empty(rows, 2, N, device="cuda")
    out = gemm(A, K, True, B, K, True, M, N, K, _st(), stats=stats)
    assert rel_l2(stats[:, 0].sum(0), out.sum(0)) < 1e-5
    assert rel_l2(stats[:, 1].sum(0), (out * out).sum(0)) < 1e-5
    assert rel_l2(out, A.double() @ B.double().t()) < 1e-5


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 16, 200, 128, 128), (3, 7, 13, 32, 64), (1, 32, 40, 64, 128), (2, 5, 9, 128, 32), (1, 1, 1, 32, 64)])
def test_implicit_gemm_conv3x3(N, H, W, cin, cout):
    from ocrs_models_b200 import _lib
    from ocrs_models_b200.rec_engine import _w_dgrad, _w_fwd, conv3x3

    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    rows = _lib.lib().ocrs_gemm_stat_rows(N * H * W)
    stats = torch.empty(rows, 2, cout, device="cuda")
    out = conv3x3(xd, N, H, W, cin, _w_fwd(w.cuda()), cout, _st(), bias=b.cuda(), stats=stats)
    assert rel_l2(out.reshape(N, H, W, cout).permute(0, 3, 1, 2), ref) < 3e-6
    assert rel_l2(stats[:, 0].sum(0), out.sum(0)) < 1e-5
    out = conv3x3(xd, N, H, W, cin, _w_fwd(w.cuda()), cout, _st(), bias=b.cuda(), relu=True)
    assert rel_l2(out.reshape(N, H, W, cout).permute(0, 3, 1, 2), torch.relu(ref)) < 3e-6
    # data gradient = the same kernel on dY with flipped / transposed weights
    dy = torch.randn(N, cout, H, W, generator=g)
    dref = F.conv_transpose2d(dy.double(), w.double(), padding=1)
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dx = conv3x3(dyd, N, H, W, cout, _w_dgrad(w.cuda()), cin, _st())
    assert rel_l2(dx.reshape(N, H, W, cin).permute(0, 3, 1, 2), dref) < 3e-6
    # weight gradient, also gathered by TMA (no im2col buffer)
    from ocrs_models_b200.grads import materialize
    from ocrs_models_b200.rec_engine import conv3x3_wgrad

    xg = x.double().requires_grad_(True)
    wg = w.double().requires_grad_(True)
    F.conv2d(xg, wg, None, padding=1).backward(dy.double())
    dwp = conv3x3_wgrad(dyd.reshape(N * H * W, cout), xd, N, H, W, cin, cout, _st())
    # split-K partial rows in the GEMM layout [cout][(ky,kx,ci)]; the multi-tensor delivery kernel reduces and re-lays them out
    assert rel_l2(materialize(dwp, w.shape, _st()), wg.grad) < 5e-6


def test_conv0_fwd_bwd():
    from ocrs_models_b200._lib import call, lib, ptr

    g = torch.Generator().manual_seed(2)
    N, H, W = 2, 64, 52
    x = torch.rand(N, 1, H, W, generator=g) - 0.5
    w = (torch.randn(32, 1, 3, 3, generator=g) * 0.5).double().requires_grad_(True)
    b = (torch.randn(32, generator=g) * 0.1).double().requires_grad_(True)
    ref = F.max_pool2d(F.relu(F.conv2d(x.double(), w, b, padding=1)), 2)
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    xd, wd, bd = x.cuda(), w.detach().float().cuda(), b.detach().float().cuda()
    out = torch.empty(N, H // 2, W // 2, 32, device="cuda")
    code = torch.empty(N * (H // 2) * (W // 2) * 4, dtype=torch.int32, device="cuda")
    call("ocrs_rec_conv0_fwd", ptr(xd), N, H, W, ptr(wd), ptr(bd), ptr(out), ptr(code), _st())
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 1e-5
    out2 = torch.empty_like(out)
    call("ocrs_rec_conv0_fwd", ptr(xd), N, H, W, ptr(wd), ptr(bd), ptr(out2), None, _st())  # inference: no code
    assert torch.equal(out, out2)
    blocks = lib().ocrs_rec_conv0_bwd_blocks()
    part = torch.empty(blocks, 32, 10, device="cuda")
    dd = dout.permute(0, 2, 3, 1).contiguous().cuda()
    call("ocrs_rec_conv0_bwd", ptr(xd), N, H, W, ptr(wd), ptr(bd), ptr(dd), None, ptr(part), _st())  # recomputing path
    got = part.sum(0)
    assert rel_l2(got[:, :9].reshape(32, 1, 3, 3), w.grad) < 1e-4
    assert rel_l2(got[:, 9], b.grad) < 1e-4
    part2 = torch.empty_like(part)
    call("ocrs_rec_conv0_bwd", ptr(xd), N, H, W, ptr(wd), ptr(bd), ptr(dd), ptr(code), ptr(part2), _st())  # saved arg-max / gates
    assert torch.equal(part, part2)


@pytest.mark.parametrize("T,N", [(9, 3), (25, 64), (6, 70)])
def test_gru_layer_fwd_bwd(T, N):
    from ocrs_models_b200._lib import call, ptr
    from ocrs_models_b200.rec_engine import gemm

    g = torch.Generator().manual_seed(T)
    I, Hd = 128, 256
    x = torch.randn(T, N, I, generator=g)
    P = {}
    for sfx in ("", "_reverse"):
        P["w_ih" + sfx] = torch.randn(3 * Hd, I, generator=g) * 0.1
        P["w_hh" + sfx] = torch.randn(3 * Hd, Hd, generator=g) * 0.1
        P["b_ih" + sfx] = torch.randn(3 * Hd, generator=g) * 0.1
        P["b_hh" + sfx] = torch.randn(3 * Hd, generator=g) * 0.1
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    x64 = x.double().requires_grad_(True)
    outs = [O._gru_direction(x64, P64["w_ih" + s], P64["w_hh" + s], P64["b_ih" + s], P64["b_hh" + s], r) for s, r in (("", False), ("_reverse", True))]
    ref = torch.cat(outs, dim=2)
    dout = torch.randn(T, N, 2 * Hd, generator=g)
    ref.backward(dout.double())

    D = {k: v.cuda() for k, v in P.items()}
    xd = x.cuda()
    st = _st()
    gi = [gemm(xd, I, True, D["w_ih" + s], I, True, T * N, 768, I, st, bias=D["b_ih" + s]) for s in ("", "_reverse")]
    out = torch.empty(T, N, 512, device="cuda")
    gates = torch.empty(T, N, 2, 4, 256, device="cuda")
    call("ocrs_gru_layer_fwd_persist", ptr(gi[0]), ptr(gi[1]), ptr(D["w_hh"]),
         ptr(D["w_hh_reverse"]), ptr(D["b_hh"]), ptr(D["b_hh_reverse"]), ptr(out), ptr(gates), T, N, st)
    assert rel_l2(out, ref) < 1e-5
    whhT = [D["w_hh" + s].t().contiguous() for s in ("", "_reverse")]
    dgi = [torch.empty(T * N, 768, device="cuda") for _ in range(2)]
    dgh = [torch.empty(T * N, 768, device="cuda") for _ in range(2)]
    dd = dout.cuda()
    call("ocrs_gru_layer_bwd_persist", ptr(whhT[0]), ptr(whhT[1]), ptr(dd), ptr(out), ptr(gates), ptr(dgi[0]),
         ptr(dgi[1]), ptr(dgh[0]), ptr(dgh[1]), T, N, st)
    for d, s in enumerate(("", "_reverse")):
        assert rel_l2(dgi[d].sum(0), P64["b_ih" + s].grad) < 1e-4
        assert rel_l2(dgh[d].sum(0), P64["b_hh" + s].grad) < 1e-4
        assert rel_l2(dgi[d].t() @ xd.reshape(T * N, I), P64["w_ih" + s].grad) < 1e-4
    dx = dgi[0] @ D["w_ih"] + dgi[1] @ D["w_ih_reverse"]
    assert rel_l2(dx.reshape(T, N, I), x64.grad) < 1e-4


def test_log_softmax():
    from ocrs_models_b200._lib import call, ptr

    g = torch.Generator().manual_seed(1)
    x = (torch.randn(77, 97, generator=g) * 4).double().requires_grad_(True)
    ref = F.log_softmax(x, dim=1)
    gr = torch.randn(77, 97, generator=g)
    ref.backward(gr.double())
    xd = x.detach().float().cuda()
    y = torch.empty_like(xd)
    call("ocrs_log_softmax_fwd", ptr(xd), ptr(y), 77, 97, _st())
    assert rel_l2(y, ref) < 1e-6
    dx = torch.empty_like(xd)
    gd = gr.cuda()
    call("ocrs_log_softmax_bwd", ptr(y), ptr(gd), ptr(dx), 77, 97, 97, _st())
    assert rel_l2(dx, x.grad) < 1e-5
    # padded row pitch (what the engine uses so that the head's gradient GEMMs can take dlog by TMA): zeros in the padding
    dxp = torch.full((77, 100), float("nan"), device="cuda")
    call("ocrs_log_softmax_bwd", ptr(y), ptr(gd), ptr(dxp), 77, 97, 100, _st())
    assert torch.equal(dxp[:, :97], dx) and dxp[:, 97:].abs().max().item() == 0


def _rec_model(seed=1234):
    from ocrs_models_b200 import RecognitionModel

    torch.manual_seed(seed)
    m = RecognitionModel(O.DEFAULT_ALPHABET)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k in ("4", "10", "16", "20"):
            m.conv[k].weight.add_(torch.randn(m.conv[k].weight.shape, generator=g) * 0.2)
            m.conv[k].bias.add_(torch.randn(m.conv[k].bias.shape, generator=g) * 0.2)
    return m


def _rec_batch(N, W, S, g, ragged=True):
    x = torch.rand(N, 1, 64, W, generator=g) - 0.5
    tgt = torch.randint(1, 97, (N, S), generator=g, dtype=torch.int32)
    tgt[0, 1] = tgt[0, 0]
    T = W // 4 + 1
    il = torch.full((N,), W // 4, dtype=torch.int64)
    tl = torch.randint(1, S + 1, (N,), generator=g) if ragged else torch.full((N,), S)
    tl[0] = S
    return {"image": x, "targets": tgt, "input_lengths": il, "target_lengths": tl}


# (5, 64, 4) is avoided on purpose: one conv.13 ReLU unit sits within fp32 rounding of 0 there and the
# flipped mask alone moves the upstream gradients by 2e-3 (a discontinuity, see SURVEY finding 6).
@pytest.mark.parametrize("N,W,S", [(3, 96, 8), (2, 200, 20), (4, 64, 4), (6, 128, 12)])
def test_full_model_train_step_vs_oracle(N, W, S):
    from ocrs_models_b200 import CTCLoss

    m = _rec_model()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    batch = _rec_batch(N, W, S, torch.Generator().manual_seed(0))
    out64, loss64, g64, nb64 = O.train_step_grads("rec", sd, batch, torch.float64)
    out32, loss32, g32, _ = O.train_step_grads("rec", sd, batch, torch.float32)
    m = m.cuda().train()
    lp = m(batch["image"].cuda())
    assert lp.shape == (W // 4 + 1, N, 97)
    loss = CTCLoss()(lp, batch["targets"].cuda(), batch["input_lengths"], batch["target_lengths"])
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(lp, out64) < 1e-4, "log-probs"
    assert abs(loss.item() - loss64.item()) < 1e-4 * abs(loss64.item())
    ours = {k: p.grad for k, p in m.named_parameters()}
    gn = torch.sqrt(sum((v.double() ** 2).sum() for v in g64.values()))
    err = torch.sqrt(sum(((ours[k].cpu().double() - g64[k]) ** 2).sum() for k in g64)) / gn
    err32 = torch.sqrt(sum(((g32[k].double() - g64[k]) ** 2).sum() for k in g64)) / gn
    print(f"global grad rel-L2: ours {err:.3e}  fp32-oracle {err32:.3e}")
    floor = gn / len(g64) ** 0.5
    bad = []
    for k in g64:
        e = (ours[k].cpu().double() - g64[k]).norm()
        e32 = (g32[k].double() - g64[k]).norm()
        if not e <= max(1e-3 * g64[k].norm(), 1e-3 * floor, 10 * e32):
            bad.append((k, float(e), float(g64[k].norm()), float(e32)))
    assert not bad, bad
    assert err < 1e-3
    for k, v in nb64.items():
        if v.is_floating_point():
            assert rel_l2(m.state_dict()[k], v) < 1e-4, k
        else:
            assert int(m.state_dict()[k]) == int(v)


def test_full_model_golden_and_eval(golden):
    from ocrs_models_b200 import CTCLoss, RecognitionModel

    gold = golden("rec_w96")
    torch.manual_seed(1234)
    m = RecognitionModel(O.DEFAULT_ALPHABET).cuda().train()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 1, 64, 96, generator=g) - 0.5
    lp = m(x.cuda())
    loss = CTCLoss()(lp, torch.from_numpy(gold["targets"]).cuda(), torch.from_numpy(gold["il"]), torch.from_numpy(gold["tl"]))
    loss.backward()
    np.testing.assert_allclose(lp.detach().cpu().numpy(), gold["lp"], rtol=1e-3, atol=1e-4)
    assert abs(loss.item() - gold["loss"]) < 1e-4 * gold["loss"]
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())).item()
    assert abs(gn - gold["grad_norm"]) < 1e-3 * gold["grad_norm"]
    m.eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.inference_mode():
        le = m(x.cuda())
    assert rel_l2(le, O.rec_forward(sd, x, training=False)) < 1e-4


def test_runs_under_autocast_like_train_rec():
    """train_rec.py:118 wraps forward+loss in autocast(bfloat16); the modules ignore it (fp32 inside)."""
    from ocrs_models_b200 import CTCLoss

    m = _rec_model().cuda().train()
    batch = _rec_batch(2, 64, 4, torch.Generator().manual_seed(1))
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        lp = m(batch["image"].cuda())
        loss = CTCLoss()(lp, batch["targets"].cuda(), batch["input_lengths"], batch["target_lengths"])
    assert lp.dtype == torch.float32 and torch.isfinite(loss)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("case", ["fwd", "dgrad", "wgrad"])
@pytest.mark.parametrize("N,cin,cout,HW", [(3, 64, 128, 1024), (2, 256, 256, 256), (4, 32, 64, 4096)])
def test_batched_gemm_on_planar_activations(case, N, cin, cout, HW):
    """ocrs_gemm_tc_batched: the three contractions of a 1x1 convolution on planar [N][C][HW] tensors (detection levels
    with >= 64 channels): y = W x (+ per-channel row statistics), g = W^T dy, dW = sum_n dy x^T; vs fp64 einsum."""
    from ocrs_models_b200 import _lib
    from ocrs_models_b200._lib import call, ptr

    lib = _lib.lib()
    g = torch.Generator().manual_seed(N * cin + HW)
    W = torch.randn(cout, cin, generator=g) * 0.1
    x = torch.randn(N, cin, HW, generator=g)
    dy = torch.randn(N, cout, HW, generator=g)
    Wd, xd, dyd = W.cuda(), x.cuda(), dy.cuda()
    st = _st()
    if case == "fwd":
        y = torch.full((N, cout + 5, HW), float("nan"), device="cuda")  # a channel slice of a wider buffer
        rows = lib.ocrs_gemm_tc_batched_stat_rows(HW, N)
        stats = torch.full((rows, 2, cout), float("nan"), device="cuda")
        call("ocrs_gemm_tc_batched", ptr(Wd), cin, 1, cout, 0, ptr(xd), HW, 0, N * cin, cin, ptr(y), HW, (cout + 5) * HW,
             cout, HW, cin, N, ptr(stats), st)
        ref = torch.einsum("oi,nip->nop", W.double(), x.double())
        assert rel_l2(y[:, :cout], ref) < 3e-6
        assert torch.isnan(y[:, cout:]).all()
        assert rel_l2(stats.double().sum(0)[0], ref.sum((0, 2))) < 1e-5 or ref.sum((0, 2)).abs().max() < 1e-2
        assert rel_l2(stats.double().sum(0)[1], (ref ** 2).sum((0, 2))) < 1e-5
    elif case == "dgrad":
        gout = torch.empty((N, cin, HW), device="cuda")
        call("ocrs_gemm_tc_batched", ptr(Wd), cin, 0, cout, 0, ptr(dyd), HW, 0, N * cout, cout, ptr(gout), HW, cin * HW,
             cin, HW, cout, N, None, st)
        assert rel_l2(gout, torch.einsum("oi,nop->nip", W.double(), dy.double())) < 3e-6
    else:
        part = torch.empty((N, cout, cin), device="cuda")
        call("ocrs_gemm_tc_batched", ptr(dyd), HW, 1, N * cout, cout, ptr(xd), HW, 1, N * cin, cin, ptr(part), cin, cout * cin,
             cout, cin, HW, N, None, st)
        # the tensor core's truncating accumulation costs ~8e-7 per 1024 accumulated products
        assert rel_l2(part.double().sum(0), torch.einsum("nop,nip->oi", dy.double(), x.double())) < 1e-6 * max(3, HW / 1024 * 1.5)
        # the same with every sample's pixel range cut into K slices (more tiles for the deep levels' small M x N)
        ks = lib.ocrs_gemm_tc_batched_splits(HW, 3)
        assert ks == (3 if HW >= 96 else 1)
        part2 = torch.full((N * ks, cout, cin), float("nan"), device="cuda")
        call("ocrs_gemm_tc_batched_splitk", ptr(dyd), HW, 1, N * cout, cout, ptr(xd), HW, 1, N * cin, cin, ptr(part2), cin,
             cout * cin, cout, cin, HW, N, 3, st)
        tol = 1e-6 * max(3, HW / 1024 * 1.5)
        assert rel_l2(part2.double().sum(0), torch.einsum("nop,nip->oi", dy.double(), x.double())) < tol
        per_sample = part2.view(N, ks, cout, cin).double().sum(1)
        assert rel_l2(per_sample, torch.einsum("nop,nip->noi", dy.double(), x.double())) < tol
