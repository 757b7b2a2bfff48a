"""C-ABI surface: the shared object loads and exports every symbol include/*.h declares.
No compute calls (there is no GPU here); without a device the entry points must fail loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "anorag_b200.h")).read()
    return sorted(set(re.findall(r"ARB_API[^;(]*?\b(arb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("arb_index_create", "arb_index_add", "arb_index_search", "arb_bm25_create", "arb_bm25_search",
                 "arb_bm25_scores", "arb_fuse", "arb_merge_topk", "arb_last_error"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    import ano_rag_b200
    from ano_rag_b200 import _native as nat
    assert os.path.exists(nat.LIB_PATH)
    lib = ctypes.CDLL(nat.LIB_PATH)
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{sym} declared in include/anorag_b200.h but not exported"
    assert set(nat.EXPORTED) == set(declared_symbols())
    assert nat.lib.arb_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import ano_rag_b200
    from ano_rag_b200 import _native as nat, engine
    with pytest.raises(nat.NativeError):
        nat.device_count()
    with pytest.raises(nat.NativeError):
        engine.DenseIndex(64)
    # the drop-in classes keep the reference's never-raise convention: log and return False / []
    from ano_rag_b200.vector_store import VectorIndex
    vi = VectorIndex(64)
    assert vi.create_index("Flat") is False
    assert vi.search([[0.0] * 64]) == []


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "ano-rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
