"""The C-ABI boundary: include/ocrs_b200.h, the ctypes signature table and the built library agree.
No compute calls: runs without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from ocrs_models_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ocrs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(ocrs_[A-Za-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        decls[m.group(2)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return decls


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_is_valid_c():
    subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", HEADER], check=True)


def test_every_declared_symbol_is_exported(lib):
    decls = _declared()
    assert len(decls) >= 50
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in ocrs_b200.h but not exported by libocrs_b200.so"


def test_ctypes_table_matches_header():
    decls = _declared()
    table = dict(_lib.SIGNATURES)
    table["ocrs_last_error"] = []
    assert set(table) == set(decls), set(table) ^ set(decls)
    for name, args in decls.items():
        assert len(args) == len(table[name]), name
        for a, ct in zip(args, table[name]):
            want = (_lib.P if "*" in a else _lib.L if a.startswith("long long") else _lib.D if a.startswith("double")
                    else _lib.F if a.startswith("float") else _lib.I)
            assert ct is want, (name, a)


def test_no_torch_types_cross_the_boundary():
    for name, args in _declared().items():
        for a in args:
            assert re.match(r"^(const )?(float|int|double|long long|void)\b", a), (name, a)


def test_host_side_queries_need_no_gpu(lib):
    assert lib.ocrs_version() >= 100
    # alpha rows hold (blank, label) pairs, NP per lane: 2 * NP * ceil((S + 1) / NP) floats
    assert lib.ocrs_ctc_alpha_row(40) == 84 and lib.ocrs_ctc_alpha_row(255) == 512 and lib.ocrs_ctc_alpha_row(256) == 0
    assert lib.ocrs_ctc_alpha_row(0) == 2 and lib.ocrs_ctc_alpha_row(31) == 64 and lib.ocrs_ctc_alpha_row(63) == 128
    assert lib.ocrs_det_dwpw_partial_rows(2, 64, 64) == 2 * 2 * 2
    assert lib.ocrs_gemm_splits(1152, 4) == 4
    assert lib.ocrs_gemm_tc_splits(1152, 5) in (4, 5)
    assert lib.ocrs_launch_count() >= 0


def test_product_path_has_no_cpu_fallback():
    import torch

    from ocrs_models_b200 import CTCLoss, DetectionModel, RecognitionModel, balanced_cross_entropy_loss
    from oracle.functional import DEFAULT_ALPHABET

    with pytest.raises(RuntimeError, match="no CPU path"):
        DetectionModel()(torch.zeros(1, 1, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        RecognitionModel(DEFAULT_ALPHABET)(torch.zeros(1, 1, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        CTCLoss()(torch.zeros(4, 1, 5), torch.zeros(1, 2, dtype=torch.int32), [4], [2])
    with pytest.raises(RuntimeError, match="no CPU path"):
        balanced_cross_entropy_loss(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ocrs_models_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
