"""Gradient delivery (ocrs_models_b200/grads.py, csrc/glue.cu): with an optim.FusedAdam attached the backward passes
accumulate straight into the flat gradient bucket; the values must equal what autograd's own accumulation of the
returned tensors gives (reference semantics: loss.backward() at train_rec.py:130 / train_detection.py:96), including
accumulation over two backward passes, and BatchNorm's num_batches_tracked must still count the training forwards."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(kind):
    from ocrs_models_b200 import CTCLoss, DetectionModel, RecognitionModel, balanced_cross_entropy_loss
    from ocrs_models_b200.alphabet import DEFAULT_ALPHABET

    g = torch.Generator().manual_seed(5)
    if kind == "rec":
        b = {"image": (torch.rand(3, 1, 64, 160, generator=g) - 0.5).cuda(), "targets": torch.randint(1, 97, (3, 12), generator=g, dtype=torch.int32).cuda(),
             "input_lengths": torch.tensor([40, 38, 40], dtype=torch.int32).cuda(), "target_lengths": torch.tensor([12, 7, 2], dtype=torch.int32).cuda()}
        ctc = CTCLoss()
        return (lambda: RecognitionModel(DEFAULT_ALPHABET)), (lambda m: ctc(m(b["image"]), b["targets"], b["input_lengths"], b["target_lengths"]))
    b = {"image": (torch.rand(2, 1, 128, 128, generator=g) - 0.5).cuda(), "mask": (torch.rand(2, 1, 128, 128, generator=g) < 0.1).float().cuda()}
    return DetectionModel, (lambda m: balanced_cross_entropy_loss(m(b["image"]), b["mask"]))


@pytest.mark.parametrize("kind", ["rec", "det"])
def test_bucket_accumulation_equals_autograd_accumulation(kind):
    from ocrs_models_b200.optim import FusedAdam

    make, loss_of = _case(kind)
    torch.manual_seed(7)
    plain = make().cuda().train()
    torch.manual_seed(7)
    sunk = make().cuda().train()
    opt = FusedAdam(sunk, lr=1e-3)
    assert all(getattr(p, "_ocrs_grad_sink", None) is not None for p in sunk.parameters())
    # same BatchNorm buffers / weights in both models
    loss_of(plain).backward()
    opt.zero_grad()
    loss_of(sunk).backward()
    for (n, a), b in zip(plain.named_parameters(), sunk.parameters()):
        assert b.grad.data_ptr() == b._ocrs_grad_sink.data_ptr(), n
        assert torch.equal(a.grad, b.grad), n
    for (n, a), b in zip(plain.named_buffers(), sunk.buffers()):
        assert torch.equal(a, b), n
        if n.endswith("num_batches_tracked"):
            assert int(a.item()) == 1
    # a second backward accumulates (the running statistics moved, so compare against the plain model doing the same)
    loss_of(plain).backward()
    loss_of(sunk).backward()
    for (n, a), b in zip(plain.named_parameters(), sunk.parameters()):
        assert torch.equal(a.grad, b.grad), n


def test_grad_deliver_maps_and_accumulate():
    """The column maps of ocrs_grad_deliver against plain torch: identity with many rows, the convolution re-layout,
    runs at a pitch with an offset; store and accumulate."""
    from ocrs_models_b200._lib import stream_ptr
    from ocrs_models_b200.grads import Partial, materialize

    st = stream_ptr(torch.device("cuda:0"))
    g = torch.Generator().manual_seed(0)
    part = torch.randn(300, 5 * 7, generator=g).cuda()
    out = materialize(Partial(part, 300, 35), (5, 7), st)
    assert torch.allclose(out, part.double().sum(0).float().view(5, 7), rtol=0, atol=1e-5)
    co, ci, khw = 6, 10, 9
    part = torch.randn(4, co * khw * ci, generator=g).cuda()
    out = materialize(Partial(part, 4, co * ci * khw, conv=(ci, khw)), (co, ci, 3, 3), st)
    ref = part.double().sum(0).view(co, 3, 3, ci).permute(0, 3, 1, 2).float()
    assert torch.equal(out, ref.contiguous())
    part = torch.randn(20, 32, 10, generator=g).cuda()
    w = materialize(Partial(part, 20, 288, ld=320, inner=(9, 10)), (32, 9), st)
    b = materialize(Partial(part, 20, 32, ld=320, off=9, inner=(1, 10)), (32,), st)
    full = part.double().sum(0).float()
    assert torch.equal(w, full[:, :9].contiguous()) and torch.equal(b, full[:, 9].contiguous())
    # wide path with 16-byte alignment, one row, accumulate
    from ocrs_models_b200._lib import call

    src = torch.randn(5000, generator=g).cuda()
    dst = torch.randn(5000, generator=g).cuda()
    want = dst + src
    one = lambda T, v: (T * 1)(v)  # noqa: E731
    call("ocrs_grad_deliver", one(ctypes.c_void_p, src.data_ptr()), one(ctypes.c_void_p, dst.data_ptr()), one(ctypes.c_int, 5000),
         one(ctypes.c_int, 1), one(ctypes.c_int, 5000), one(ctypes.c_int, 0), one(ctypes.c_int, 1), one(ctypes.c_int, 1),
         one(ctypes.c_int, 1), 1, st)
    assert torch.equal(dst, want)


def test_weight_prep_layouts():
    from ocrs_models_b200 import RecognitionModel
    from ocrs_models_b200._lib import stream_ptr
    from ocrs_models_b200.alphabet import DEFAULT_ALPHABET
    from ocrs_models_b200.rec_engine import Split, _w_dgrad, _w_fwd, prepare_weights

    dev = torch.device("cuda:0")
    m = RecognitionModel(DEFAULT_ALPHABET).cuda()
    prep = prepare_weights(m, stream_ptr(dev), dev, True)
    for name in ("3", "9", "19"):
        w = m.conv[name].weight
        for key, ref in (("fwd", _w_fwd(w)), ("dg", _w_dgrad(w))):
            got = prep[key][id(w)]
            assert isinstance(got, Split)
            assert torch.equal(got.hi + got.lo, ref), (name, key)
            assert (got.hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0  # hi is TF32-exact
    w = m.gru.weight_hh_l1_reverse
    assert torch.equal(prep["whhT"][id(w)], w.detach().t().contiguous())
    w = m.gru.weight_ih_l0
    s = prep["w"][id(w)]
    assert torch.equal(s.hi + s.lo, w.detach())
