"""TEST INFRASTRUCTURE (oracle): proof wire format, restating bulletproofs 4.0.0 `R1CSProof::to_bytes` /
`from_bytes` (src/r1cs/proof.rs; crate pinned at Cargo.lock:47-50, sources not vendored - restated from the published
crate, unpinned: the reference itself serialises nothing) for the one-phase case:

    to_bytes:   [0u8] | A_I1 A_O1 S1 | T_1 T_3 T_4 T_5 T_6 | t_x t_x_blinding e_blinding | ipp (L_i R_i pairs, a, b)
    from_bytes: empty slice, unknown version byte, (len - 1) % 32 != 0, too few elements, an inner-product
                part that is not 2 lg n points + 2 scalars, or a non-canonical scalar -> ProofError::FormatError
                (here the element-count checks collapse into "the record has exactly the circuit's length");
                points are kept compressed (decompression failures surface in verify as VerificationError).

Modes 0 / 1 of this repo (l and r in the clear) have no upstream format; they use the version bytes 0x80 / 0x81
with the same rules (every field after the 8 points is a scalar).  Only tests/ may import this module."""
from . import ristretto255 as R

VERSION = {0: 0x80, 1: 0x81, 2: 0x00}


def next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


def proof_len(n, mode):
    if mode == 2:
        return 32 * (13 + 2 * (next_pow2(n).bit_length() - 1))
    return 32 * (11 + 2 * n)


def to_bytes(proof: bytes, mode: int) -> bytes:
    return bytes([VERSION[mode]]) + proof


def from_bytes(rec: bytes, n: int, mode: int):
    """-> proof bytes, or None for ProofError::FormatError."""
    if len(rec) == 0 or rec[0] != VERSION[mode]:
        return None
    body = rec[1:]
    if len(body) % 32 or len(body) != proof_len(n, mode):
        return None
    words = len(body) // 32
    if mode == 2:
        lg = (words - 13) // 2
        scalar_idx = [8, 9, 10] + list(range(11 + 2 * lg, words))
    else:
        scalar_idx = range(8, words)
    for i in scalar_idx:
        if int.from_bytes(body[32 * i:32 * i + 32], "little") >= R.L:
            return None
    return body
