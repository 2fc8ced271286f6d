"""Device-side batch assembly (csrc/data.cu, ocrs_models_b200/data.py) vs the UNMODIFIED reference collate_samples +
transform_image: bit exact, same keys, same dropped samples."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_loader  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not staged")]


def _samples(g, u8):
    out = []
    for i, (w, s) in enumerate([(120, 9), (800, 40), (256, 64), (37, 3), (500, 20), (24, 7)]):
        img = torch.randint(0, 256, (1, 64, w), generator=g, dtype=torch.uint8)
        seq = torch.randint(1, 97, (s,), generator=g, dtype=torch.int32)
        if i == 5:
            seq[:] = 5  # 7 equal labels need 13 frames but 24 // 4 = 6: dropped by ctc_input_and_target_compatible
        out.append({"image": img, "text_seq": seq})
    return out


@pytest.mark.parametrize("u8", [True, False])
def test_collate_matches_reference(u8):
    ref_loader.load()
    from ocrs_models.datasets.util import transform_image
    from ocrs_models.train_rec import collate_samples as ref_collate

    from ocrs_models_b200.data import collate_samples

    g = torch.Generator().manual_seed(3)
    raw = _samples(g, u8)
    ref_in = [{"image": transform_image(s["image"]), "text_seq": s["text_seq"].clone()} for s in raw]
    ours_in = [{"image": s["image"] if u8 else transform_image(s["image"]), "text_seq": s["text_seq"].clone()} for s in raw]
    ref = ref_collate(ref_in)
    got = collate_samples(ours_in, "cuda")
    assert ref["image"].shape == got["image"].shape == (5, 1, 64, 1024)  # 800 -> 1024, the incompatible sample is dropped
    assert torch.equal(got["image"].cpu(), ref["image"])
    assert got["text_seq"].dtype == ref["text_seq"].dtype and torch.equal(got["text_seq"].cpu(), ref["text_seq"])
    assert torch.equal(got["text_len"], ref["text_len"]) and torch.equal(got["image_width"], ref["image_width"])
    for k in [k for k in sys.modules if k == "ocrs_models" or k.startswith("ocrs_models.")]:
        del sys.modules[k]
