"""CPU test: libalva_b200.so loads without a GPU and exports every symbol include/alva_b200.h declares;
creating a context without a device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "alva_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(alva_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol():
    import alvaar_b200
    L = C.CDLL(alvaar_b200.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/alva_b200.h but not exported: {missing}"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import alvaar_b200
    with pytest.raises(alvaar_b200.AlvaError):
        alvaar_b200.Context(0)


def test_product_does_not_touch_oracle():
    """The product path may not import / link / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "alvaar_b200")
    for dp, _, files in os.walk(pkg):
        if "_build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "libalva_oracle" not in txt and "libalva_ref" not in txt and "orc_" not in txt, (dp, f)
