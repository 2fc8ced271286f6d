"""The CUDA-graph training step (optim.GraphedTrainStep) must follow the eager step exactly: same losses, same weights,
same BatchNorm buffers and Adam state after several steps (the replay runs the very kernels the capture recorded; the
step count for Adam's bias correction lives on the device)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["rec", "det"])
def test_graphed_step_equals_eager_step(kind):
    from ocrs_models_b200 import CTCLoss, DetectionModel, RecognitionModel, balanced_cross_entropy_loss
    from ocrs_models_b200.alphabet import DEFAULT_ALPHABET
    from ocrs_models_b200.optim import FusedAdam, GraphedTrainStep

    g = torch.Generator().manual_seed(2)
    if kind == "rec":
        batches = [{"image": (torch.rand(3, 1, 64, 128, generator=g) - 0.5).cuda(), "targets": torch.randint(1, 97, (3, 16), generator=g, dtype=torch.int32).cuda(),
                    "input_lengths": torch.tensor([32, 30, 32], dtype=torch.int32).cuda(), "target_lengths": torch.tensor([16, 7, 2], dtype=torch.int32).cuda()}
                   for _ in range(4)]
        ctc = CTCLoss()
        loss_of = lambda m, b: ctc(m(b["image"]), b["targets"], b["input_lengths"], b["target_lengths"])  # noqa: E731
        make = lambda: RecognitionModel(DEFAULT_ALPHABET)  # noqa: E731
        clip = 4.0
    else:
        batches = [{"image": (torch.rand(2, 1, 96, 64, generator=g) - 0.5).cuda(), "mask": (torch.rand(2, 1, 96, 64, generator=g) < 0.1).float().cuda()}
                   for _ in range(4)]
        loss_of = lambda m, b: balanced_cross_entropy_loss(m(b["image"]), b["mask"])  # noqa: E731
        make = DetectionModel
        clip = None
    results = []
    for graphed in (False, True):
        torch.manual_seed(1234)
        model = make().cuda().train()
        opt = FusedAdam(model, lr=1e-3, max_grad_norm=clip)
        losses = []
        if graphed:
            sd0 = {k: v.clone() for k, v in model.state_dict().items()}
            step = GraphedTrainStep(model, opt, loss_of, batches[0], warmup=2)
            # the warm-up and the capture pass ran real steps: rewind model, optimiser moments and step counter
            model.load_state_dict(sd0)
            opt.m.zero_(); opt.v.zero_(); opt.step_dev.zero_(); opt.t = 0
            for b in batches:
                losses.append(float(step(b).item()))
        else:
            for b in batches:
                opt.zero_grad()
                loss = loss_of(model, b)
                loss.backward()
                opt.step()
                losses.append(float(loss.item()))
        results.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}, int(opt.step_dev.item())))
    (l0, s0, t0), (l1, s1, t1) = results
    assert t0 == t1 == 4
    assert l0 == l1, (l0, l1)
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k


def test_prefetched_batches_give_the_same_steps():
    """GraphedTrainStep.prefetch: batches copied on the copy stream one step ahead must produce exactly the steps that
    `step(batch)` produces (same losses, same weights)."""
    from ocrs_models_b200 import DetectionModel, balanced_cross_entropy_loss
    from ocrs_models_b200.optim import FusedAdam, GraphedTrainStep

    g = torch.Generator().manual_seed(3)
    batches = [{"image": (torch.rand(2, 1, 64, 64, generator=g) - 0.5).pin_memory(), "mask": (torch.rand(2, 1, 64, 64, generator=g) < 0.1).float().pin_memory()}
               for _ in range(5)]
    loss_of = lambda m, b: balanced_cross_entropy_loss(m(b["image"]), b["mask"])  # noqa: E731
    out = []
    for prefetch in (False, True):
        torch.manual_seed(1234)
        model = DetectionModel().cuda().train()
        opt = FusedAdam(model, lr=1e-3)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        step = GraphedTrainStep(model, opt, loss_of, batches[0], warmup=2)
        model.load_state_dict(sd0)
        opt.m.zero_(); opt.v.zero_(); opt.step_dev.zero_(); opt.t = 0
        losses = []
        if prefetch:
            step.prefetch(batches[0])
            for i in range(len(batches)):
                loss = step()
                if i + 1 < len(batches):
                    step.prefetch(batches[i + 1])  # overlaps the replay just launched
                losses.append(float(loss.item()))
        else:
            for b in batches:
                losses.append(float(step(b).item()))
        out.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert out[0][0] == out[1][0]
    for k in out[0][1]:
        assert torch.equal(out[0][1][k], out[1][1][k]), k
