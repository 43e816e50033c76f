"""CPU: the C-ABI library loads without a GPU and exports every symbol include/b200dsp.h
declares; the ctypes table in newsched_b200 covers exactly that set; no product code touches
the oracle."""
import ctypes
import os
import re

import pytest

import newsched_b200 as nb
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert len(names) >= 60
    for must in ("b200_copy", "b200_fir_run", "b200_fft_run", "b200_pfb_run", "b200_ring_create",
                 "b200_chain_run_host"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(nb.LIB_PATH), "libb200dsp.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(nb.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(L, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert sorted(nb.SIGNATURES) == _declared()


def test_version_and_error_text_without_gpu():
    L = nb.lib()
    assert L.b200_version() >= 100
    assert isinstance(L.b200_last_error(), bytes)
    # argument errors are reported, not crashed on, even with no device
    assert L.b200_fir_create(None, None) == -1
    assert b"fir_create" in L.b200_last_error()
    assert L.b200_fft_create(None, None) == -1


def test_no_cpu_fallback():
    import torch
    with pytest.raises(nb.B200Error):
        nb.copy(torch.zeros(4))
    with pytest.raises(nb.B200Error):
        nb.multiply_const(torch.zeros(4, dtype=torch.complex64), 2.0)


def test_product_never_imports_oracle():
    bad = []
    for base in ("newsched_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|liboracle|oracle/|orc_)", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
