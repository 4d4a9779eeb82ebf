"""Host-side mirror of the reference's ACProof module (circuit_lib.rs) on the CUDA backend.

`Circuit` + `Generators` are ACEssentials (circuit_lib.rs:58-74); a `Batch` carries `count` ACProver
states (a_L, a_R, a_O, gamma: circuit_lib.rs:76-81) through the reference's seven steps at once.
Scalars are ints or 32-byte little-endian bytes; points are 32-byte compressed encodings.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

from .backend import Backend, scalars_to_bytes

MODE_REFERENCE, MODE_REFERENCE_FIXED, MODE_FIXED = 0, 1, 2
_MODES = {"reference": 0, "reference-fixed": 1, "fixed": 2, 0: 0, 1: 1, 2: 2}


def _sc32(x) -> bytes:
    return x.to_bytes(32, "little") if isinstance(x, int) else bytes(x)


class Circuit:
    """W_L, W_R, W_O (n x Q), W_V (m x Q) as (wire, constraint, coeff) triples and c (Q)."""

    def __init__(self, backend: Backend, n: int, Q: int, m: int, WL, WR, WO, WV, c_vec):
        self.be, self.n, self.Q, self.m = backend, n, Q, m
        mats = [list(WL), list(WR), list(WO), list(WV)]
        nnz = (ctypes.c_uint32 * 4)(*[len(t) for t in mats])
        flat = [t for mat in mats for t in mat]
        N = max(len(flat), 1)
        wire = (ctypes.c_uint32 * N)(*[t[0] for t in flat])
        cons = (ctypes.c_uint32 * N)(*[t[1] for t in flat])
        coeff = b"".join(_sc32(t[2]) for t in flat) or bytes(32)
        if len(c_vec) != Q:
            raise ValueError("c_vec must have Q entries")
        h = ctypes.c_void_p()
        backend._check(backend._lib.bpp_circuit_create(backend._ctx, n, Q, m, nnz, wire, cons, coeff,
                                                        scalars_to_bytes(c_vec), ctypes.byref(h)))
        self._h = h
        backend._adopt(self)

    @classmethod
    def shuffle(cls, backend: Backend, k: int):
        """The corrected k-card shuffle circuit, built inside the library (bpp_circuit_create_shuffle)."""
        self = cls.__new__(cls)
        self.be, self.n, self.Q, self.m = backend, 2 * k, 4 * k, 2 * k + 1
        h = ctypes.c_void_p()
        backend._check(backend._lib.bpp_circuit_create_shuffle(backend._ctx, k, ctypes.byref(h)))
        self._h = h
        backend._adopt(self)
        return self

    @classmethod
    def from_dense(cls, backend, W_L, W_R, W_O, W_V, c_vec):
        """From the reference's dense matrices (rows = wires, columns = constraints)."""
        def trip(M):
            return [(i, q, v) for i, row in enumerate(M) for q, v in enumerate(row) if v]
        return cls(backend, len(W_L), len(W_L[0]), len(W_V), trip(W_L), trip(W_R), trip(W_O), trip(W_V), c_vec)

    def free(self):
        if self._h and self.be._ctx:
            self.be._lib.bpp_circuit_free(self.be._ctx, self._h)
        self._h = None


class Generators:
    """g_base, h_base, G_vec, H_vec + their fixed-base window tables on the device."""

    def __init__(self, backend: Backend, g: bytes, h: bytes, G: Sequence[bytes], H: Sequence[bytes], window_bits: int = 0):
        if len(G) != len(H):
            raise ValueError("G_vec and H_vec must have the same length (circuit_lib.rs:154)")
        self.be, self.n = backend, len(G)
        hnd = ctypes.c_void_p()
        backend._check(backend._lib.bpp_gens_create(backend._ctx, g, h, b"".join(G), b"".join(H), len(G), window_bits,
                                                     ctypes.byref(hnd)))
        self._h = hnd
        backend._adopt(self)

    def free(self):
        if self._h and self.be._ctx:
            self.be._lib.bpp_gens_free(self.be._ctx, self._h)
        self._h = None


def next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def proof_len(n: int, mode="reference-fixed") -> int:
    """Bytes per proof: modes "reference"/"reference-fixed" carry l and r in the clear (circuit_lib.rs:464-468),
    "fixed" replaces them by the inner-product proof (2 lg n' points + 2 scalars)."""
    if _MODES[mode] == 2:
        return 32 * (13 + 2 * (next_pow2(n).bit_length() - 1))
    return 32 * (11 + 2 * n)


class Batch:
    """`count` proofs in lock-step; buffers stay on the device between calls."""

    def __init__(self, backend: Backend, circuit: Circuit, gens: Generators, count: int, mode="reference-fixed",
                 label: bytes = b"test"):
        self.be, self.cir, self.gens, self.count, self.mode = backend, circuit, gens, count, _MODES[mode]
        h = ctypes.c_void_p()
        backend._check(backend._lib.bpp_acp_batch_create(backend._ctx, circuit._h, gens._h, self.mode, count, label,
                                                          len(label), ctypes.byref(h)))
        self._h = h
        backend._adopt(self)
        self.proof_len = proof_len(circuit.n, self.mode)

    def upload_witness(self, aL: bytes, aR: bytes, aO: bytes, gamma: bytes, seeds: bytes):
        n, m, B = self.cir.n, self.cir.m, self.count
        if not (len(aL) == len(aR) == len(aO) == 32 * n * B and len(gamma) == 32 * m * B and len(seeds) == 32 * B):
            raise ValueError("witness shape mismatch (circuit_lib.rs:160-167 assert_eq!)")
        self.be._check(self.be._lib.bpp_acp_batch_upload_witness(self._h, aL, aR, aO, gamma, seeds))

    def gen_shuffle_witness(self, deck, perm, x, gamma, seeds):
        """Witness of `count` k-card shuffles computed on the device: deck (k x 32 bytes), perm (count x k uint32 indices),
        x (count x 32), gamma (count x m x 32), seeds (count x 32); bytes or addresses of pinned buffers."""
        self.be._check(self.be._lib.bpp_acp_batch_gen_shuffle_witness(self._h, deck, perm, x, gamma, seeds))

    def commit(self, v: bytes = None, want: bool = True) -> bytes:
        out = ctypes.create_string_buffer(32 * self.cir.m * self.count) if want else None
        self.be._check(self.be._lib.bpp_acp_batch_commit(self._h, v, out))
        return out.raw if want else b""

    def upload_commitments(self, V):
        """The value commitments (count x m x 32 compressed; bytes or the address of a pinned buffer): modes
        "reference-fixed" and "fixed" bind them to every proof's transcript, so the prover needs them before prove()."""
        if not isinstance(V, int) and len(V) != 32 * self.cir.m * self.count:
            raise ValueError("V: count x m x 32 bytes")
        self.be._check(self.be._lib.bpp_acp_batch_upload_commitments(self._h, V))

    def set_host_transcripts(self, on: bool):
        """Fiat-Shamir on host threads (True) instead of one device thread per proof (default)."""
        self.be._check(self.be._lib.bpp_acp_batch_set_host_transcripts(self._h, 1 if on else 0))

    def set_batch_rlc(self, on: bool):
        """Verify with one random-linear-combination MSM over the batch first (default) or per proof only."""
        self.be._check(self.be._lib.bpp_acp_batch_set_batch_rlc(self._h, 1 if on else 0))

    def set_priority_split(self, on: bool):
        """Table-gather MSMs on an internal lowest-priority stream (for several batches in flight on urgent streams)."""
        self.be._check(self.be._lib.bpp_acp_batch_set_priority_split(self._h, 1 if on else 0))

    def prove(self):
        self.be._check(self.be._lib.bpp_acp_batch_prove(self._h))

    # raw-address forms for callers that own pinned host buffers (addresses as ints)
    def upload_witness_ptr(self, aL: int, aR: int, aO: int, gamma: int, seeds: int):
        self.be._check(self.be._lib.bpp_acp_batch_upload_witness(self._h, aL, aR, aO, gamma, seeds))

    def download_proofs_ptr(self, dst: int):
        self.be._check(self.be._lib.bpp_acp_batch_download_proofs(self._h, dst))

    def upload_proofs_ptr(self, proofs: int, V: int = None):
        self.be._check(self.be._lib.bpp_acp_batch_upload_proofs(self._h, proofs, V))

    def download_accept_ptr(self, dst: int):
        self.be._check(self.be._lib.bpp_acp_batch_download_accept(self._h, dst))

    def download_proofs(self) -> bytes:
        out = ctypes.create_string_buffer(self.proof_len * self.count)
        self.be._check(self.be._lib.bpp_acp_batch_download_proofs(self._h, out))
        return out.raw

    def upload_proofs(self, proofs: bytes, V: bytes = None):
        self.be._check(self.be._lib.bpp_acp_batch_upload_proofs(self._h, proofs, V))

    def verify(self, verifier_seed: bytes = None):
        """verifier_seed: 32 bytes of secret, fresh randomness; None = the library draws them from the OS.  The
        per-proof and batch weights are challenge scalars of each proof's transcript rekeyed with it."""
        if verifier_seed is not None and len(verifier_seed) != 32:
            raise ValueError("verifier_seed: 32 bytes or None")
        self.be._check(self.be._lib.bpp_acp_batch_verify(self._h, verifier_seed))

    def download_accept(self) -> bytes:
        out = ctypes.create_string_buffer(self.count)
        self.be._check(self.be._lib.bpp_acp_batch_download_accept(self._h, out))
        return out.raw

    def gather_accept(self, per: int) -> bytes:
        """Sharded batch verification: all ranks' accept bytes (nranks x per, rank r's at r * per), through the
        library's NCCL communicator (Backend.comm_init)."""
        nr, _ = self.be.comm_info()
        out = ctypes.create_string_buffer(per * nr)
        self.be._check(self.be._lib.bpp_acp_batch_gather_accept(self._h, per, out))
        return out.raw

    def time_commit_msm(self, reps: int = 5):
        ms, madd, add = ctypes.c_float(), ctypes.c_uint64(), ctypes.c_uint64()
        self.be._check(self.be._lib.bpp_acp_batch_time_commit_msm(self._h, reps, ctypes.byref(ms), ctypes.byref(madd),
                                                                  ctypes.byref(add)))
        return ms.value, madd.value, add.value

    def free(self):
        if self._h and self.be._ctx:
            self.be._lib.bpp_acp_batch_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def prove_batch(backend, circuit, gens, aL, aR, aO, gamma, seeds, count, mode="reference-fixed", label=b"test",
                V: bytes = None) -> bytes:
    """V: the value commitments (count x m x 32 compressed) - required in every mode but "reference"."""
    if _MODES[mode] != 0 and V is None:
        raise ValueError("modes reference-fixed and fixed bind the commitments V to the transcript: pass V")
    out = ctypes.create_string_buffer(proof_len(circuit.n, mode) * count)
    backend._check(backend._lib.bpp_acproof_prove_batch(backend._ctx, circuit._h, gens._h, _MODES[mode], count, aL, aR, aO,
                                                         gamma, seeds, V, label, len(label), out))
    return out.raw


def verify_batch(backend, circuit, gens, proofs, V, count, mode="reference-fixed", label=b"test",
                 verifier_seed=None) -> bytes:
    out = ctypes.create_string_buffer(count)
    backend._check(backend._lib.bpp_acproof_verify_batch(backend._ctx, circuit._h, gens._h, _MODES[mode], count, proofs, V,
                                                          label, len(label), verifier_seed, out))
    return out.raw


def transcript_script(backend, records) -> bytes:
    """Run a merlin::Transcript on the device.  records: ("append", label, message) or ("challenge", label, n);
    the first record must be ("append", b"dom-sep", protocol_label) == Transcript::new(protocol_label).
    Returns the concatenated challenge bytes."""
    script, out_len = b"", 0
    for op, label, arg in records:
        if op == "append":
            script += bytes([0, len(label)]) + label + len(arg).to_bytes(4, "little") + arg
        else:
            script += bytes([1, len(label)]) + label + int(arg).to_bytes(4, "little")
            out_len += int(arg)
    out = ctypes.create_string_buffer(max(out_len, 1))
    backend._check(backend._lib.bpp_transcript_script(backend._ctx, script, len(script), out, out_len))
    return out.raw[:out_len]


# ---- proof wire format: byte handling only, no device and no Backend needed ---------------
WIRE_VERSION = {0: 0x80, 1: 0x81, 2: 0x00}   # mode 2 = bulletproofs 4.0.0 R1CSProof::to_bytes, one-phase


def wire_len(n: int, mode="fixed") -> int:
    from . import _lib
    return _lib.load().bpp_acproof_wire_len(n, _MODES[mode])


def to_wire(proofs: bytes, n: int, count: int, mode="fixed") -> bytes:
    """count proofs (the library's proof bytes) -> count wire records (version byte + proof)."""
    from . import _lib
    lib = _lib.load()
    if len(proofs) != count * proof_len(n, mode):
        raise ValueError("proofs: wrong length for this circuit")
    out = ctypes.create_string_buffer(count * wire_len(n, mode))
    rc = lib.bpp_acproof_to_wire(n, _MODES[mode], count, proofs, out)
    if rc:
        raise _lib.BppError(rc, lib.bpp_strerror(rc).decode())
    return out.raw


def from_wire(wire: bytes, n: int, count: int, mode="fixed"):
    """-> (proof bytes, status bytes): status 1 = ProofError::FormatError (wrong version byte or a non-canonical
    scalar; that proof's bytes are zeroed).  Raises on a record length that does not match the circuit."""
    from . import _lib
    lib = _lib.load()
    if count == 0 or len(wire) % count:
        raise _lib.BppError(-4, lib.bpp_strerror(-4).decode())
    out = ctypes.create_string_buffer(count * proof_len(n, mode))
    status = ctypes.create_string_buffer(count)
    rc = lib.bpp_acproof_from_wire(n, _MODES[mode], count, wire, len(wire) // count, out, status)
    if rc:
        raise _lib.BppError(rc, lib.bpp_strerror(rc).decode())
    return out.raw, status.raw
