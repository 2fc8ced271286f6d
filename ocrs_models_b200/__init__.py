"""B200-native (sm_100a) training hot paths of robertknight/ocrs-models.

Drop-in surface (reference file:line in each module's docstring):

* ``DetectionModel``, ``RecognitionModel``  <- ocrs_models/models.py
* ``balanced_cross_entropy_loss``           <- ocrs_models/train_detection.py:225-263
* ``CTCLoss``                               <- torch.nn.CTCLoss as used by ocrs_models/train_rec.py:104
* ``RecognitionAccuracyStats``              <- ocrs_models/train_rec.py:20-82 (greedy decode + CER, on the device)
* ``install()`` rebinds those names inside an imported ``ocrs_models`` so that its unmodified
  ``train_detection.py`` / ``train_rec.py`` run on the CUDA kernels.
"""
from .losses import CTCLoss, balanced_cross_entropy_loss  # noqa: F401
from .metrics import RecognitionAccuracyStats, greedy_decode_cer  # noqa: F401
from .models import DetectionModel, RecognitionModel  # noqa: F401

__all__ = ["DetectionModel", "RecognitionModel", "CTCLoss", "balanced_cross_entropy_loss", "RecognitionAccuracyStats",
           "greedy_decode_cer", "install"]


def install(package: str = "ocrs_models") -> list[str]:
    """Rebind the hot-path names inside an importable reference package so that its UNMODIFIED
    scripts run on these modules: ``models.DetectionModel`` / ``models.RecognitionModel``
    (ocrs_models/models.py:93,146), the globals ``DetectionModel`` and
    ``balanced_cross_entropy_loss`` that ``train_detection.main`` resolves
    (train_detection.py:376,419,446,454) and ``RecognitionModel`` / ``CTCLoss`` that ``train_rec``
    resolves (train_rec.py:104,180,379), plus ``RecognitionAccuracyStats`` (train_rec.py:100,176) so the per-batch
    greedy decode + edit distance run on the device. Returns the list of rebound names."""
    import importlib

    done = []
    plan = {
        f"{package}.models": {"DetectionModel": DetectionModel, "RecognitionModel": RecognitionModel},
        f"{package}.train_detection": {"DetectionModel": DetectionModel,
                                       "balanced_cross_entropy_loss": balanced_cross_entropy_loss},
        f"{package}.train_rec": {"RecognitionModel": RecognitionModel, "CTCLoss": CTCLoss,
                                 "RecognitionAccuracyStats": RecognitionAccuracyStats},
        f"{package}.eval_detection": {"DetectionModel": DetectionModel},
    }
    for modname, names in plan.items():
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        for name, obj in names.items():
            if hasattr(mod, name):
                setattr(mod, name, obj)
                done.append(f"{modname}.{name}")
    if not done:
        raise ImportError(f"could not import {package}: nothing rebound")
    return done
