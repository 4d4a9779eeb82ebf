"""include/bpperm.h through a C11 compiler and a plain C host (tests/c/abi_smoke.c): the header is C, the entry points
have the types it declares, and a C program drives MSM + prove + verify to the oracle's bytes."""
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, SRC, "-ldl"])
    return exe


def test_header_is_c11_and_symbols_have_the_declared_types(tmp_path):
    import bpperm_b200
    exe = _build(tmp_path)
    out = subprocess.check_output([exe, bpperm_b200._lib.LIB_PATH, "symbols"], text=True)
    assert "symbols ok" in out


@pytest.mark.gpu
def test_c_host_runs_msm_and_a_proof_batch_to_the_oracles_bytes(tmp_path):
    import bpperm_b200
    from oracle import acproof as A, ipa, ristretto255 as R
    from oracle.chacha import ChaChaRng
    exe = _build(tmp_path)
    sb = lambda v: b"".join(R.sc_bytes(s) for s in v)
    # the golden MSM vector of the GPU tests
    rng = ChaChaRng(bytes(range(32)))
    pts = [rng.point() for _ in range(4)]
    sc = [rng.scalar() for _ in range(4)]
    blob = struct.pack("<I", 4) + sb(sc) + b"".join(R.compress(p) for p in pts)
    # one `fixed`-mode batch of a 3-card shuffle
    k, mode, count = 3, 2, 2
    core, prover, V = ipa.make_instance(k, ChaChaRng(b"\x21" * 32), dense_weights=True)
    n, Q, m = core["n"], core["Q"], core["m"]
    mats = core["sparse"]
    flat = [t for M in mats for t in M]
    ng = len(core["G_vec"])
    seeds = [b"\x31" * 32, b"\x32" * 32]
    Vc = b"".join(R.compress(p) for p in V)
    blob += struct.pack("<9I", n, Q, m, mode, count, *[len(M) for M in mats])
    blob += b"".join(struct.pack("<I", t[0]) for t in flat) + b"".join(struct.pack("<I", t[1]) for t in flat)
    blob += b"".join(R.sc_bytes(t[2]) for t in flat) + sb(core["c_vec"])
    blob += R.compress(core["g_base"]) + R.compress(core["h_base"]) + struct.pack("<I", ng)
    blob += b"".join(R.compress(p) for p in core["G_vec"]) + b"".join(R.compress(p) for p in core["H_vec"])
    blob += sb(prover["a_L"]) * count + sb(prover["a_R"]) * count + sb(prover["a_O"]) * count + sb(prover["gamma"]) * count
    blob += b"".join(seeds) + Vc * count
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.write_bytes(blob)
    res = subprocess.run([exe, bpperm_b200._lib.LIB_PATH, "run", str(inp), str(outp)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = outp.read_bytes()
    assert out[:32].hex() == "ac0188282e26885b30102aa5ee4e91734b3328dba689b17245af361585d58d6f"
    plen = ipa.proof_len(n)
    for i, sd in enumerate(seeds):
        want, _ = ipa.prove(core, prover, V, ChaChaRng(sd))
        assert out[32 + i * plen:32 + (i + 1) * plen] == want
    assert out[32 + count * plen:] == b"\x01" * count
