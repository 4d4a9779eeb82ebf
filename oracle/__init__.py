"""CPU oracle for the bulletproof-perm hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic the reference
(ercembu/bulletproof-perm) performs on its hot path: Ristretto255 group
operations and multiscalar multiplication (curve25519-dalek-ng 4.1.1), scalar
arithmetic mod l, the Merlin transcript (merlin 3.0.0) and the reference's own
operator/protocol code (bp-perm/src/{util,poly,transcript_protocol,circuit_lib,
weights}.rs).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import or execute anything under `oracle/`, and
there only as the checker, never as the product.  The product path
(`bulletproof-perm_b200/`) never imports this package and fails loudly when the
CUDA library is missing.

Parity pins (the reference itself has no golden vectors - SURVEY.md section 4):
RFC 9496 ristretto255 vectors, the published Merlin KAT, libsodium 1.0.20
cross-check fixtures in tests/golden/, and the SURVEY.md appendix E vectors.
The third-party crates that hold the algorithms (curve25519-dalek-ng 4.1.1,
merlin 3.0.0, bulletproofs 4.0.0, rand_chacha 0.3.1) are NOT vendored in
/root/reference; their published algorithms are restated here.
"""
