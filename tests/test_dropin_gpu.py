"""Drop-in behaviour: the loops of train_detection.py:82-111 and train_rec.py:107-153, restated here with the
stock torch optimiser / clipping / autocast, driving our modules; trajectories vs the CPU oracle."""
import pytest
import torch

from oracle import functional as O

pytestmark = pytest.mark.gpu


def _oracle_traj(kind, sd, batch, steps, clip):
    sd = {k: v.clone() for k, v in sd.items()}
    state, losses = {}, []
    for _ in range(steps):
        _, loss, grads, nb = O.train_step_grads(kind, sd, batch)
        if clip:
            O.clip_grad_norm(grads, clip)
        O.adam_step({k: sd[k] for k in grads}, grads, state)
        sd.update(nb)
        losses.append(float(loss))
    return losses, sd


def test_detection_loop_like_train_detection():
    from ocrs_models_b200 import DetectionModel, balanced_cross_entropy_loss

    torch.manual_seed(1234)
    model = DetectionModel()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2, 1, 128, 96, generator=g) - 0.5
    masks = (torch.rand(2, 1, 128, 96, generator=g) < 0.1).float()
    ref_losses, ref_sd = _oracle_traj("det", sd0, {"image": img, "mask": masks}, 3, None)
    device = torch.device("cuda")
    model = model.to(device)
    optimizer = torch.optim.Adam(model.parameters())
    model.train()
    losses = []
    for _ in range(3):
        pred_masks = model(img.to(device))
        loss = balanced_cross_entropy_loss(pred_masks, masks.to(device))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    assert set(model.state_dict()) == set(ref_sd)
    for k, v in model.state_dict().items():
        if v.is_floating_point() and "running" in k:
            assert torch.allclose(v.cpu(), ref_sd[k], rtol=2e-3, atol=1e-5), k


def test_recognition_loop_like_train_rec():
    from ocrs_models_b200 import CTCLoss, RecognitionModel

    torch.manual_seed(1234)
    model = RecognitionModel(O.DEFAULT_ALPHABET)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    img = torch.rand(4, 1, 64, 256, generator=g) - 0.5  # collate_samples pads widths to multiples of 256
    text_seq = torch.randint(1, 97, (4, 64), generator=g, dtype=torch.int32)
    image_width = torch.tensor([200, 256, 180, 120])
    input_lengths = image_width.div(4, rounding_mode="floor")
    target_lengths = torch.tensor([20, 30, 11, 3])
    batch = {"image": img, "targets": text_seq, "input_lengths": input_lengths, "target_lengths": target_lengths}
    ref_losses, _ = _oracle_traj("rec", sd0, batch, 3, 4.0)
    device = torch.device("cuda")
    model = model.to(device).train()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = CTCLoss()
    losses = []
    for _ in range(3):
        optimizer.zero_grad()
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            pred_seq = model(img.to(device))
            batch_loss = loss_fn(pred_seq, text_seq.to(device), input_lengths, target_lengths)
        assert pred_seq.shape == (65, 4, 97)
        _ = pred_seq[:, 0, :].argmax(-1)[: input_lengths[0]].tolist()  # stats.update / preview path
        batch_loss.backward()
        grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=4.0)
        assert torch.isfinite(grad_norm)
        optimizer.step()
        losses.append(batch_loss.item())
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)


def test_fused_step_equals_stock_step():
    from ocrs_models_b200 import CTCLoss, RecognitionModel
    from ocrs_models_b200.optim import FusedAdam

    g = torch.Generator().manual_seed(3)
    img = (torch.rand(2, 1, 64, 64, generator=g) - 0.5).cuda()
    tgt = torch.randint(1, 97, (2, 4), generator=g, dtype=torch.int32).cuda()
    il, tl = torch.tensor([16, 16]), torch.tensor([4, 2])
    outs = []
    for fused in (False, True):
        torch.manual_seed(1234)
        m = RecognitionModel(O.DEFAULT_ALPHABET).cuda().train()
        opt = FusedAdam(m, lr=1e-3, max_grad_norm=4.0) if fused else torch.optim.Adam(m.parameters(), lr=1e-3)
        for _ in range(2):
            opt.zero_grad()
            CTCLoss()(m(img), tgt, il, tl).backward()
            if not fused:
                torch.nn.utils.clip_grad_norm_(m.parameters(), 4.0)
            opt.step()
        outs.append({k: v.detach().clone() for k, v in m.state_dict().items()})
    for k in outs[0]:
        if outs[0][k].is_floating_point():
            assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-4, atol=1e-6), k
