"""Generates tests/golden/libsodium_ristretto255.json with libsodium 1.0.20 (an implementation
independent of both the oracle and the CUDA code; bundled with pyzmq in this image).
Run once in the build container; the JSON is committed, the GPU box never runs this."""
import ctypes
import glob
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.chacha import ChaChaRng  # only used as a deterministic byte source

cands = glob.glob("/opt/prime-rl/.venv/lib/python3.12/site-packages/pyzmq.libs/libsodium*.so*")
sod = ctypes.CDLL(cands[0])
assert sod.sodium_init() >= 0


def from_hash(b64):
    out = ctypes.create_string_buffer(32)
    assert sod.crypto_core_ristretto255_from_hash(out, b64) == 0
    return out.raw


def scalarmult(s, p):
    out = ctypes.create_string_buffer(32)
    rc = sod.crypto_scalarmult_ristretto255(out, s, p)
    return out.raw if rc == 0 else bytes(32)  # libsodium returns -1 for the identity result


def add(p, q):
    out = ctypes.create_string_buffer(32)
    assert sod.crypto_core_ristretto255_add(out, p, q) == 0
    return out.raw


def reduce64(b):
    out = ctypes.create_string_buffer(32)
    sod.crypto_core_ristretto255_scalar_reduce(out, b)
    return out.raw


def sc_mul(a, b):
    out = ctypes.create_string_buffer(32)
    sod.crypto_core_ristretto255_scalar_mul(out, a, b)
    return out.raw


def sc_add(a, b):
    out = ctypes.create_string_buffer(32)
    sod.crypto_core_ristretto255_scalar_add(out, a, b)
    return out.raw


def sc_inv(a):
    out = ctypes.create_string_buffer(32)
    assert sod.crypto_core_ristretto255_scalar_invert(out, a) == 0
    return out.raw


rng = ChaChaRng(hashlib.sha256(b"bpperm golden v1").digest())
vec = {"source": "libsodium 1.0.20 (pyzmq bundled)", "from_hash": [], "scalar_reduce": [], "scalarmult": [], "add": [],
       "scalar_ops": [], "msm": []}
pts = []
for _ in range(48):
    b = rng.fill_bytes(64)
    e = from_hash(b)
    pts.append(e)
    vec["from_hash"].append({"in": b.hex(), "out": e.hex()})
scalars = []
for _ in range(48):
    b = rng.fill_bytes(64)
    s = reduce64(b)
    scalars.append(s)
    vec["scalar_reduce"].append({"in": b.hex(), "out": s.hex()})
for i in range(24):
    vec["scalarmult"].append({"s": scalars[i].hex(), "p": pts[i].hex(), "out": scalarmult(scalars[i], pts[i]).hex()})
    vec["add"].append({"p": pts[i].hex(), "q": pts[i + 24].hex(), "out": add(pts[i], pts[i + 24]).hex()})
    a, b = scalars[i], scalars[i + 24]
    vec["scalar_ops"].append({"a": a.hex(), "b": b.hex(), "mul": sc_mul(a, b).hex(), "add": sc_add(a, b).hex(),
                              "inv_a": sc_inv(a).hex()})
for n in (1, 2, 5, 16, 48):
    acc = None
    for i in range(n):
        t = scalarmult(scalars[i], pts[i])
        acc = t if acc is None else add(acc, t)
    vec["msm"].append({"n": n, "out": acc.hex()})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsodium_ristretto255.json")
json.dump(vec, open(out, "w"), indent=0)
print("wrote", out)
