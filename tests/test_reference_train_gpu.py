"""Row n2 of the coverage table: the reference's UNMODIFIED `train_detection.train()` and `train_rec.train()`
(staged under baseline/_ref by scripts/stage_reference.py) drive our CUDA modules after `ocrs_models_b200.install()`,
with a synthetic Dataset/DataLoader; the loss trajectories are compared with the CPU oracle's."""
import os
import sys

import pytest
import torch
from torch.utils.data import DataLoader, Dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402
from oracle import functional as O  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not staged")]


@pytest.fixture()
def installed():
    import ocrs_models_b200 as ours

    ref_loader.load()
    done = ours.install("ocrs_models")
    yield ours, done
    for k in [k for k in sys.modules if k == "ocrs_models" or k.startswith("ocrs_models.")]:
        del sys.modules[k]


class _DetData(Dataset):
    def __init__(self, n, h, w):
        g = torch.Generator().manual_seed(3)
        self.img = torch.rand(n, 1, h, w, generator=g) - 0.5
        self.mask = (torch.rand(n, 1, h, w, generator=g) < 0.1).float()

    def __len__(self):
        return self.img.shape[0]

    def __getitem__(self, i):
        return {"path": f"synthetic-{i}", "image": self.img[i], "text_mask": self.mask[i]}


class _RecData(Dataset):
    def __init__(self, n):
        g = torch.Generator().manual_seed(4)
        self.items = []
        for i in range(n):
            w = int(torch.randint(120, 500, (1,), generator=g))
            s = int(torch.randint(3, 20, (1,), generator=g))
            self.items.append((torch.rand(1, 64, w, generator=g) - 0.5, torch.randint(1, 97, (s,), generator=g, dtype=torch.int32)))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        img, seq = self.items[i]
        return {"image": img.clone(), "text_seq": seq.clone()}


def _oracle_mean_loss(kind, sd, batches, clip):
    sd = {k: v.clone() for k, v in sd.items()}
    state, losses = {}, []
    for b in batches:
        _, loss, grads, nb = O.train_step_grads(kind, sd, b)
        if clip:
            O.clip_grad_norm(grads, clip)
        O.adam_step({k: sd[k] for k in grads}, grads, state)
        sd.update(nb)
        losses.append(float(loss))
    return sum(losses) / len(losses), sd


def test_unmodified_train_detection_runs_on_our_modules(installed):
    ours, done = installed
    from ocrs_models import train_detection

    assert "ocrs_models.train_detection.balanced_cross_entropy_loss" in done
    assert train_detection.train.__module__ == "ocrs_models.train_detection"  # the reference's own loop
    torch.manual_seed(1234)
    model = train_detection.DetectionModel()
    assert isinstance(model, ours.DetectionModel)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    data = _DetData(4, 160, 120)
    dl = DataLoader(data, batch_size=2, shuffle=False)
    device = torch.device("cuda")
    model = model.to(device)
    opt = torch.optim.Adam(model.parameters())
    mean_loss = train_detection.train(0, device, dl, model, train_detection.balanced_cross_entropy_loss, opt)
    batches = [{"image": data.img[i:i + 2], "mask": data.mask[i:i + 2]} for i in (0, 2)]
    ref_loss, ref_sd = _oracle_mean_loss("det", sd0, batches, None)
    print(f"train_detection.train(): mean loss {mean_loss:.6f}, oracle {ref_loss:.6f}")
    assert abs(mean_loss - ref_loss) < 2e-3 * abs(ref_loss)
    assert int(model.state_dict()["in_conv.seq.0.seq.2.num_batches_tracked"]) == 2


def test_unmodified_train_rec_runs_on_our_modules(installed):
    ours, done = installed
    from ocrs_models import train_rec

    assert "ocrs_models.train_rec.CTCLoss" in done
    assert train_rec.train.__module__ == "ocrs_models.train_rec"
    torch.manual_seed(1234)
    model = train_rec.RecognitionModel(alphabet=train_rec.DEFAULT_ALPHABET)
    assert isinstance(model, ours.RecognitionModel)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    data = _RecData(8)
    dl = DataLoader(data, batch_size=4, shuffle=False, collate_fn=train_rec.collate_samples)
    device = torch.device("cuda")
    model = model.to(device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    mean_loss, stats = train_rec.train(0, device, dl, model, opt)
    batches = []
    for b in DataLoader(data, batch_size=4, shuffle=False, collate_fn=train_rec.collate_samples):
        assert b["image"].shape[-1] % 256 == 0  # collate pads to the next multiple of 256 (train_rec.py:267-269)
        batches.append({"image": b["image"], "targets": b["text_seq"], "input_lengths": b["image_width"].div(4, rounding_mode="floor"),
                        "target_lengths": b["text_len"]})
    ref_loss, _ = _oracle_mean_loss("rec", sd0, batches, 4.0)
    print(f"train_rec.train(): mean loss {mean_loss:.6f}, oracle {ref_loss:.6f}; CER {stats.char_error_rate():.4f}")
    assert abs(mean_loss - ref_loss) < 2e-3 * abs(ref_loss)
    assert 0.0 < float(stats.char_error_rate()) < 10.0 and int(stats.total_chars) == sum(int(b["target_lengths"].sum()) for b in batches)
