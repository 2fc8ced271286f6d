"""Input pipeline of the recognition path on the device.

:func:`collate_samples` is a drop-in for reference ``ocrs_models/train_rec.py:248-304`` when training on CUDA: same
filtering rule (``ctc_input_and_target_compatible``, train_rec.py:220-239), same padded sizes (``round_up`` adds a full
unit when already aligned, :242-245,267-272), same keys and dtypes in the returned dict - but the image batch is
assembled by one kernel from ONE packed host buffer (raw ``uint8`` pixels are normalised on the fly exactly like
``transform_image``, datasets/util.py:27-35), instead of N ``F.pad`` calls + a stack on the host followed by a copy of
the padded fp32 batch (4x the bytes of the raw pixels, plus the padding).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr

DOWNSAMPLE = 4          # train_rec.py:259
IMG_WIDTH_STEP = 256    # train_rec.py:265


def round_up(val: int, unit: int) -> int:
    """train_rec.py:242-245: NOT a ceiling - an aligned value still grows by one unit."""
    return val + (unit - val % unit)


def ctc_input_and_target_compatible(input_len: int, target) -> bool:
    """train_rec.py:220-239."""
    t = target.tolist() if isinstance(target, torch.Tensor) else list(target)
    need = max(1, len(t)) + sum(1 for i in range(1, len(t)) if t[i - 1] == t[i])
    return input_len >= need


def collate_samples(samples: list[dict], device=None) -> dict:
    """samples: dicts with "image" ([1, H, w] or [H, w]; uint8 raw pixels or float already transformed) and "text_seq"
    (1-D int labels). Returns {"image" [N,1,H,Wpad] f32 on `device`, "text_seq" [N,Spad] int32 on `device`,
    "text_len" [N] int64 (host), "image_width" [N] int64 (host)} like the reference's default_collate output."""
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("ocrs_models_b200.data.collate_samples has no CPU path (use the reference collate_samples)")
    widths_all = [int(s["image"].shape[-1]) for s in samples]
    max_w = round_up(max(widths_all), IMG_WIDTH_STEP)
    max_s = round_up(max(int(s["text_seq"].shape[0]) for s in samples), IMG_WIDTH_STEP // DOWNSAMPLE)
    keep = [s for s, w in zip(samples, widths_all) if ctc_input_and_target_compatible(w // DOWNSAMPLE, s["text_seq"])]
    if not keep:
        raise RuntimeError("no sample of the batch is compatible with the CTC loss")
    N = len(keep)
    imgs = [s["image"].reshape(s["image"].shape[-2], s["image"].shape[-1]) for s in keep]
    H = int(imgs[0].shape[0])
    is_u8 = imgs[0].dtype == torch.uint8
    if any(i.shape[0] != H or (i.dtype == torch.uint8) != is_u8 for i in imgs):
        raise RuntimeError("all line images of a batch must share height and dtype")
    widths = [int(i.shape[1]) for i in imgs]
    packed = torch.cat([i.reshape(-1) if is_u8 else i.float().reshape(-1) for i in imgs])
    offs, acc = [], 0
    for w in widths:
        offs.append(acc)
        acc += H * w
    text = torch.zeros((N, max_s), dtype=torch.int32)  # text_pad_value = 0, the CTC blank (train_rec.py:287)
    lens = []
    for n, s in enumerate(keep):
        L = int(s["text_seq"].shape[0])
        text[n, :L] = s["text_seq"].to(torch.int32)
        lens.append(L)
    packed_d = packed.pin_memory().to(device, non_blocking=True)
    offs_d = torch.tensor(offs, dtype=torch.int64).pin_memory().to(device, non_blocking=True)
    w_d = torch.tensor(widths, dtype=torch.int32).pin_memory().to(device, non_blocking=True)
    out = torch.empty((N, 1, H, max_w), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        call("ocrs_collate_lines", ptr(packed_d), int(is_u8), ptr(offs_d), ptr(w_d), N, H, max_w, ptr(out), _lib.stream_ptr(device))
    return {"image": out, "text_seq": text.pin_memory().to(device, non_blocking=True), "text_len": torch.tensor(lens, dtype=torch.int64),
            "image_width": torch.tensor(widths, dtype=torch.int64)}
