"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/okge_b200.h declares; the ctypes table mirrors the header one to one. No compute calls."""
import ctypes
import os
import re

import pytest

from open_knowledge_graph_embeddings_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "okge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(okge_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    assert len(syms) == 53
    for must in ("okge_gather_pool_fwd", "okge_gather_pool_bwd", "okge_fold_query", "okge_score_store", "okge_score_bce",
                 "okge_score_lse", "okge_score_rank", "okge_rank_count", "okge_adagrad_dense", "okge_adam_dense",
                 "okge_gemm_adagrad", "okge_row_slots_build"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in okge_b200.h but not exported"
    assert set(declared_symbols()) == set(_capi.SIGNATURES), "ctypes table and header disagree"
    lib.okge_abi_version.restype = ctypes.c_int
    assert lib.okge_abi_version() == _capi.ABI_VERSION


def test_no_cpu_fallback_in_product_path():
    """Ops refuse CPU tensors instead of silently computing on the host."""
    import torch
    from open_knowledge_graph_embeddings_b200 import kernels as K
    with pytest.raises(_capi.OkgeNativeError):
        K.gather_rows(torch.zeros(4, 8), torch.zeros(2, dtype=torch.int32))
    with pytest.raises(_capi.OkgeNativeError):
        K.score_store(torch.zeros(4, 8), torch.zeros(6, 8))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "open_knowledge_graph_embeddings_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
