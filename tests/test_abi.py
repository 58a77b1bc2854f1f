"""CPU: the C-ABI library loads and exports every symbol include/sky_search.h declares.
No compute calls are made here (no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sky_search.h")).read()
    return sorted(set(re.findall(r"SKY_API\s+[\w\s\*]+?\b(sky_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from sky_embeddings_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 15
    assert sorted(_lib.SIGNATURES) == decl


def test_library_exports_every_symbol():
    import __graft_entry__ as entry
    entry.build()
    from sky_embeddings_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.sky_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from sky_embeddings_b200 import Bank
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Bank(10, 1, 8)
    from sky_embeddings_b200 import similarity as S
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.compute_similarity(torch.zeros(3, 1, 8), torch.zeros(4, 1, 8), metric="cosine")
    with pytest.raises(RuntimeError):
        S.mae_simsearch(None, torch.zeros(1, 2, 8), [], "cpu", nested_batches=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sky_embeddings_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                # docstrings may CITE reference files; nothing may open / import them at run time
                assert not re.search(r"(open|load|import_module|spec_from_file_location|sys\.path\.\w+)\([^)]*/root/reference", src), f
