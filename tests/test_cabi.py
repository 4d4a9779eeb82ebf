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


def test_tools_and_bench_product_legs_do_not_import_the_oracle():
    """tools/ are product-side utilities: only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/"""
    tools = os.path.join(ROOT, "tools")
    for f in os.listdir(tools):
        if f.endswith(".py"):
            txt = open(os.path.join(tools, f)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    ours = bench[bench.index("def run_ours"):bench.index("def main")]
    # inside the GPU arm the oracle may only appear in the cpu_baseline calls
    assert "from oracle" not in ours and "import oracle" not in ours


def test_null_arguments_are_rejected_without_a_device():
    """entry points validate their arguments before touching CUDA: BPP_ERR_INVALID_ARG (-3), never a crash"""
    import ctypes
    import bpperm_b200
    lib = bpperm_b200.load()
    assert lib.bpp_acp_batch_set_host_transcripts(None, 1) == -3
    assert lib.bpp_acp_batch_set_batch_rlc(None, 0) == -3
    assert lib.bpp_transcript_script(None, b"x", 1, None, 0) == -3
    assert lib.bpp_acp_batch_prove(None) == -3
    assert lib.bpp_acp_batch_verify(None, bytes(32)) == -3
    assert lib.bpp_acp_batch_verify(None, None) == -3
    assert lib.bpp_msm_vartime(None, b"", 0, None, 0, 0, None, None) == -3
    assert lib.bpp_strerror(-3).decode() == "invalid argument"
    assert lib.bpp_acproof_proof_len_mode(104, 2) == 32 * (13 + 2 * 7)
    assert lib.bpp_acproof_proof_len_mode(104, 1) == 32 * (11 + 208)


def _header_functions():
    """name -> parameter count, parsed from include/bpperm.h (comments stripped)."""
    import re
    txt = open(os.path.join(ROOT, "include", "bpperm.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(bpp_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_every_header_function_is_exported_and_listed():
    import bpperm_b200
    lib = bpperm_b200.load()
    fns = _header_functions()
    assert len(fns) > 80
    missing = [f for f in fns if not hasattr(lib, f)]
    assert not missing, missing
    unlisted = [f for f in fns if f not in bpperm_b200.SYMBOLS]
    assert not unlisted, unlisted


def test_rust_extern_block_agrees_with_the_header():
    """rust/bpperm-sys is never compiled here; at least every `pub fn` it declares exists in the header with the same
    number of parameters."""
    import re
    fns = _header_functions()
    src = open(os.path.join(ROOT, "rust", "bpperm-sys", "src", "lib.rs")).read()
    src = re.sub(r"//[^\n]*", "", src)
    decl = re.findall(r"pub fn (bpp_\w+)\s*\(([^)]*)\)", src, flags=re.S)
    assert len(decl) > 60
    for name, args in decl:
        assert name in fns, name
        n = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n == fns[name], (name, n, fns[name])
