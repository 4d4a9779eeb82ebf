"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/bpperm.h declares, and refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bpperm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpp_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import bpperm_b200
    lib = bpperm_b200.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bpperm.h but not exported"
    assert sorted(bpperm_b200.SYMBOLS) == declared, "python binding list and header disagree"


def test_header_cites_reference_call_sites():
    src = open(os.path.join(ROOT, "include", "bpperm.h")).read()
    for needle in ("circuit_lib.rs", "vartime_multiscalar_mul", "util.rs"):
        assert needle in src


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bpperm_b200
    with pytest.raises(bpperm_b200.BppError) as ei:
        bpperm_b200.Backend(0)
    assert ei.value.status == -1


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "bulletproof-perm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
