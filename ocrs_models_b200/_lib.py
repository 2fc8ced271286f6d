"""ctypes binding of the C-ABI library (include/ocrs_b200.h).

The product path has no CPU or PyTorch fallback: if ``libocrs_b200.so`` is missing, or a call
returns non-zero, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libocrs_b200.so")

P, I, F, D, L = c_void_p, c_int, c_float, c_double, c_longlong

# name -> argtypes (all return int unless listed in _RESTYPE)
SIGNATURES: dict[str, list] = {
    "ocrs_version": [],
    "ocrs_device_arch": [],
    "ocrs_ctc_alpha_row": [I],
    "ocrs_ctc_fwd": [P, P, I, P, P, I, I, I, I, I, I, I, P, P, P, P],
    "ocrs_ctc_bwd": [P, P, I, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
}
_RESTYPE = {"ocrs_last_error": c_char_p}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libocrs_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: run `make -C {CSRC}` (or __graft_entry__.build()). "
                "There is no CPU fallback for this path."
            )
        l = ctypes.CDLL(LIB_PATH)
        l.ocrs_last_error.restype = c_char_p
        l.ocrs_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = l
    return _lib


def call(name: str, *args) -> None:
    """Invoke an entry point; raise RuntimeError with ocrs_last_error() on failure."""
    l = lib()
    rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {l.ocrs_last_error().decode()}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
