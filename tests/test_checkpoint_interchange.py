"""checkpoints interchange with the reference through unchanged state_dict keys. Uses the UNMODIFIED
reference classes and its own save_checkpoint / load_checkpoint (train_detection.py:198-215) staged under baseline/_ref;
no GPU needed (parameters only)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not staged")


@pytest.fixture(scope="module")
def ref():
    ref_loader.load()
    from ocrs_models import models, train_detection
    from ocrs_models.datasets.hiertext import DEFAULT_ALPHABET

    yield models, train_detection, DEFAULT_ALPHABET
    for k in [k for k in sys.modules if k == "ocrs_models" or k.startswith("ocrs_models.")]:
        del sys.modules[k]


def _pairs(ref):
    import ocrs_models_b200 as ours

    models, _, alphabet = ref
    return [(models.DetectionModel, ours.DetectionModel, ()), (models.RecognitionModel, ours.RecognitionModel, (alphabet,))]


def test_state_dicts_load_strictly_both_ways(ref):
    for ref_cls, our_cls, args in _pairs(ref):
        torch.manual_seed(7)
        r = ref_cls(*args)
        torch.manual_seed(8)
        o = our_cls(*args)
        assert list(r.state_dict()) == list(o.state_dict())  # same keys, same order
        o.load_state_dict(r.state_dict(), strict=True)
        for (k, a), (_, b) in zip(r.state_dict().items(), o.state_dict().items()):
            assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b), k
        torch.manual_seed(9)
        r2 = ref_cls(*args)
        r2.load_state_dict(o.state_dict(), strict=True)
        assert all(torch.equal(a, b) for a, b in zip(r.state_dict().values(), r2.state_dict().values()))


def test_reference_checkpoint_functions_round_trip_our_modules(ref, tmp_path):
    """save_checkpoint written from OUR module + stock Adam loads into the REFERENCE class (and back) with the
    reference's own functions; the optimizer state (per-parameter exp_avg / exp_avg_sq, same parameter order) survives."""
    _, td, _ = ref
    for ref_cls, our_cls, args in _pairs(ref):
        torch.manual_seed(1234)
        ours_m = our_cls(*args)
        opt = torch.optim.Adam(ours_m.parameters(), lr=1e-3)
        for p in ours_m.parameters():  # a fake step so that the optimizer has state
            p.grad = torch.full_like(p, 0.01)
        opt.step()
        path = str(tmp_path / f"{our_cls.__name__}.pt")
        td.save_checkpoint(path, ours_m, opt, epoch=3)
        ref_m = ref_cls(*args)
        ref_opt = torch.optim.Adam(ref_m.parameters(), lr=1e-3)
        ck = td.load_checkpoint(path, ref_m, ref_opt, torch.device("cpu"))
        assert ck["epoch"] == 3
        for (k, a), (_, b) in zip(ours_m.state_dict().items(), ref_m.state_dict().items()):
            assert torch.equal(a, b), k
        s_ours, s_ref = opt.state_dict()["state"], ref_opt.state_dict()["state"]
        assert s_ours.keys() == s_ref.keys()
        for i in s_ours:
            assert torch.equal(s_ours[i]["exp_avg"], s_ref[i]["exp_avg"])
        # and the reverse: a reference checkpoint resumes in our module
        path2 = str(tmp_path / f"{ref_cls.__name__}_ref.pt")
        td.save_checkpoint(path2, ref_m, ref_opt, epoch=4)
        fresh = our_cls(*args)
        fresh_opt = torch.optim.Adam(fresh.parameters(), lr=1e-3)
        assert td.load_checkpoint(path2, fresh, fresh_opt, torch.device("cpu"))["epoch"] == 4
        assert all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), ref_m.state_dict().values()))
