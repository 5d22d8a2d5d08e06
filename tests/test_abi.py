"""The C-ABI boundary: libagnn.so builds, loads, and exports every symbol that
include/agnn.h declares.  No compute calls here (no GPU needed): argument
validation happens on the host before any CUDA call."""
import ctypes as C
import os
import re

import pytest

from analysisgnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "agnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agnn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("agnn_version", "agnn_last_error", "agnn_csr_build", "agnn_csr_build_workspace",
                 "agnn_gather_reduce", "agnn_rowscale_sum", "agnn_hgt_attn_fwd", "agnn_hgt_attn_bwd_dst",
                 "agnn_hgt_attn_bwd_src"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} is declared in include/agnn.h but not exported"


def test_python_binding_covers_every_declared_symbol():
    assert sorted(_lib.exported_symbols()) == declared_symbols()


def test_version_and_error_text(lib):
    assert lib.agnn_version() >= 100
    rc = lib.agnn_gather_reduce(4, 0, 0, 0, 0, 1, None, None, 0, None, 0, 0, None, 0, None, None, 0, None)
    assert rc == -1
    assert b"gather_reduce" in lib.agnn_last_error()


def test_struct_layouts_match_the_header():
    # agnn_coo_t: 3 pointers + int64 + 4 int32 + 2 int64; agnn_rel_t: 3 pointers + int64 + pointer + 2 int32
    assert C.sizeof(_lib.Coo) == 3 * 8 + 8 + 4 * 4 + 5 * 8
    assert C.sizeof(_lib.Rel) == 3 * 8 + 8 + 8 + 2 * 4 + 3 * 8
    assert C.sizeof(_lib.HgtRel) == 6 * 8 + 8 + 2 * 8 + 8 + 2 * 4


def test_csr_workspace_is_host_arithmetic(lib):
    seg = (_lib.Coo * 1)()
    seg[0].n_edges, seg[0].n_rows, seg[0].n_cols, seg[0].n_rel = 1000, 100, 100, 4
    assert lib.agnn_csr_build_workspace(1, seg) > 0
    seg[0].n_rel = 0
    assert lib.agnn_csr_build_workspace(1, seg) == 0
    assert b"n_rel" in lib.agnn_last_error()


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors."""
    import torch
    from analysisgnn_b200 import nn as ann
    conv = ann.SageConvScatter(8, 8)
    x = torch.randn(5, 8)
    ei = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(_lib.AgnnError):
        conv(x, ei)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "analysisgnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
