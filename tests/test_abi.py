"""CPU-only: the C-ABI library loads and exports every symbol include/mfmg_b200.h declares, the ctypes table
covers all of them, the product never references oracle/, and without a GPU the product fails loudly."""
import ctypes
import os
import re

import pytest

from mfmg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mfmg_b200.h")).read()
    return sorted(set(re.findall(r"MFMGB_API\s+[\w\s\*]+?\b(mfmgb_\w+)\s*\(", text)))


def test_header_declares_a_sane_number_of_entry_points():
    syms = _declared_symbols()
    assert len(syms) >= 50
    for must in ["mfmgb_spmv", "mfmgb_residual_neg", "mfmgb_jacobi_apply", "mfmgb_restrict", "mfmgb_prolong_correct",
                 "mfmgb_dense_factor", "mfmgb_dense_solve", "mfmgb_mf_apply", "mfmgb_vcycle", "mfmgb_vcycle_host",
                 "mfmgb_pcg", "mfmgb_csr_adopt_device"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/mfmg_b200.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_version_string():
    assert b"sm_100a" in _lib.load().mfmgb_version()


def test_product_never_touches_the_oracle():
    for base in ("mfmg_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                    assert "libmfmg_oracle" not in text or f == "build.py", f


def test_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mfmg_b200.device import CudaHandle, MfmgError

    with pytest.raises(MfmgError):
        CudaHandle(0)
