"""B200-native (sm_100a) training hot paths of robertknight/ocrs-models.

Drop-in surface (reference file:line in each module's docstring):

* ``DetectionModel``, ``RecognitionModel``  <- ocrs_models/models.py
* ``balanced_cross_entropy_loss``           <- ocrs_models/train_detection.py:225-263
* ``CTCLoss``                               <- torch.nn.CTCLoss as used by ocrs_models/train_rec.py:104
* ``install()`` rebinds those names inside an imported ``ocrs_models`` so that its unmodified
  ``train_detection.py`` / ``train_rec.py`` run on the CUDA kernels.
"""
from .losses import CTCLoss, balanced_cross_entropy_loss  # noqa: F401
from .models import DetectionModel, RecognitionModel  # noqa: F401

__all__ = ["DetectionModel", "RecognitionModel", "CTCLoss", "balanced_cross_entropy_loss"]
