"""``nn.Module`` surface of the reference's ``ocrs_models/models.py`` for the two hot paths.

``DetectionModel()`` and ``RecognitionModel(alphabet)`` keep the reference's constructor
signatures, ``forward`` contracts and ``state_dict`` names/shapes (models.py:93-143,146-268), so
checkpoints interchange and ``train_detection.py`` / ``train_rec.py`` run unchanged. The module
tree below is only a *parameter container* (built in the reference's construction order, so
``torch.manual_seed(s)`` yields identical initial weights); every FLOP of ``forward`` and
``backward`` runs in the hand-written sm_100a kernels behind the C-ABI (``det_engine`` /
``rec_engine``). There is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import torch
from torch import nn

DEPTH_SCALE = [8, 16, 32, 32, 64, 128, 256]  # models.py:112


def _holder(**children) -> nn.Module:
    m = nn.Module()
    for k, v in children.items():
        m.add_module(k, v)
    return m


def _separable(cin: int, cout: int) -> nn.Module:
    # parameter names of reference DepthwiseConv (models.py:11-25): seq.0 dw3x3, seq.1 1x1, seq.2 BN
    return _holder(
        seq=nn.Sequential(
            nn.Conv2d(cin, cin, kernel_size=3, padding=1, bias=False, groups=cin),
            nn.Conv2d(cin, cout, kernel_size=1, bias=False),
            nn.BatchNorm2d(cout),
        )
    )


def _double(cin: int, cout: int) -> nn.Module:
    return _holder(seq=nn.Sequential(_separable(cin, cout), _separable(cout, cout)))


class _Validated(nn.Module):
    """Parameters/buffers are checked (device, dtype, contiguity) on the first forward after any `_apply`
    (.to / .cuda / .half / memory_format changes all go through it), not on every call."""

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_ocrs_checked", None)
        return super()._apply(fn, *args, **kwargs)


class DetectionModel(_Validated):
    """Text-detection U-Net (depthwise-separable), greyscale NCHW in, text probability out."""

    def __init__(self):
        super().__init__()
        d = DEPTH_SCALE
        self.depth_scale = d
        self.in_conv = _double(1, d[0])
        self.down = nn.ModuleList(_holder(seq=nn.Sequential(_double(d[i], d[i + 1]))) for i in range(len(d) - 1))
        self.up = nn.ModuleList()
        for i in range(len(d) - 1):
            up = nn.ConvTranspose2d(d[i + 1], d[i], kernel_size=3, stride=2)
            self.up.append(_holder(up=up, contract=_double(2 * d[i], d[i])))
        self.out_conv = nn.Sequential(nn.Conv2d(d[0], 1, kernel_size=1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .det_engine import detection_forward

        return detection_forward(self, x)


class RecognitionModel(_Validated):
    """CRNN: conv backbone -> 2-layer BiGRU -> linear -> log-softmax; (W//4+1, N, classes) out."""

    def __init__(self, alphabet: str):
        super().__init__()
        n_classes = len(alphabet) + 1
        conv = nn.ModuleDict()
        # indices follow the reference nn.Sequential (models.py:179-243); parameter-free entries omitted
        conv["0"] = nn.Conv2d(1, 32, kernel_size=3, padding=1)
        conv["3"] = nn.Conv2d(32, 64, kernel_size=3, padding=1, bias=False)
        conv["4"] = nn.BatchNorm2d(64)
        conv["7"] = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        conv["9"] = nn.Conv2d(128, 128, kernel_size=3, padding=1, bias=False)
        conv["10"] = nn.BatchNorm2d(128)
        conv["13"] = nn.Conv2d(128, 128, kernel_size=3, padding=1)
        conv["15"] = nn.Conv2d(128, 128, kernel_size=3, padding=1, bias=False)
        conv["16"] = nn.BatchNorm2d(128)
        conv["19"] = nn.Conv2d(128, 128, kernel_size=2, padding=1, bias=False)
        conv["20"] = nn.BatchNorm2d(128)
        self.conv = conv
        self.gru = _GRUParams(128, 256)
        self.output = nn.Sequential(nn.Linear(512, n_classes))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .rec_engine import recognition_forward

        return recognition_forward(self, x)


class _GRUParams(nn.Module):
    """Parameter container with nn.GRU's names, shapes, order and init (models.py:245)."""

    def __init__(self, input_size: int, hidden: int, num_layers: int = 2):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden, num_layers
        names = []
        for layer in range(num_layers):
            isz = input_size if layer == 0 else 2 * hidden
            for suffix in ("", "_reverse"):
                shapes = {
                    f"weight_ih_l{layer}{suffix}": (3 * hidden, isz),
                    f"weight_hh_l{layer}{suffix}": (3 * hidden, hidden),
                    f"bias_ih_l{layer}{suffix}": (3 * hidden,),
                    f"bias_hh_l{layer}{suffix}": (3 * hidden,),
                }
                for k, shp in shapes.items():
                    self.register_parameter(k, nn.Parameter(torch.empty(shp)))
                    names.append(k)
        stdv = 1.0 / hidden**0.5
        for k in names:  # same order and distribution as nn.RNNBase.reset_parameters
            nn.init.uniform_(getattr(self, k), -stdv, stdv)
