"""Locate and import the UNMODIFIED reference package for the reference arm of bench.py and the drop-in tests.

The reference is a pure-Python poetry project; `pip install --target baseline/_ref /root/reference` fails in this
image (build backend `poetry-core` is not installed and there is no network), so `scripts/stage_reference.py` stages
the package directory byte for byte into the git-ignored `baseline/_ref/` (which does travel to the GPU box).
Nothing of this repo's product path is imported here.
"""
from __future__ import annotations

import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "ocrs_models", "models.py"))


def load():
    """Returns the imported `ocrs_models` package (models, train_detection, train_rec importable)."""
    if not available():
        raise ImportError(f"{REF_DIR}/ocrs_models is missing: run `python scripts/stage_reference.py` in the build "
                          "container (it needs /root/reference)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for mod in ("shapely", "pylev"):
        try:
            importlib.import_module(mod)
        except ImportError:
            if STUBS not in sys.path:
                sys.path.append(STUBS)
    return importlib.import_module("ocrs_models")
