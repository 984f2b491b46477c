"""f128 helpers on Python ints (host-side parameter derivation only, never the hot path).

winter-math 0.13.1 `fields::f128::BaseElement`: p = 2^128 - 45*2^40 + 1, GENERATOR = 3,
TWO_ADICITY = 40 (reference imports: zk-lisp-proof-winterfell/src/prove.rs:20,26).
"""
P = 2**128 - 45 * 2**40 + 1
GENERATOR = 3
TWO_ADICITY = 40
TWO_ADIC_ROOT = pow(GENERATOR, (P - 1) >> TWO_ADICITY, P)
POW2_64 = 1 << 64  # utils.rs:34-45


def inv(x: int) -> int:
    return pow(x, P - 2, P) if x % P else 0


def root_of_unity(log_n: int) -> int:
    """StarkField::get_root_of_unity(log_n)."""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(TWO_ADIC_ROOT, 1 << (TWO_ADICITY - log_n), P)


def fe_to_bytes(x: int) -> bytes:
    return (x % P).to_bytes(16, "little")


def fe_from_bytes(b: bytes) -> int:
    return int.from_bytes(b[:16], "little")


def be_from_le8(b32: bytes) -> int:
    """utils.rs:346-355."""
    lo = int.from_bytes(b32[0:8], "little")
    hi = int.from_bytes(b32[8:16], "little")
    return (lo + hi * POW2_64) % P


def fold_bytes32_to_fe(b32: bytes) -> int:
    """utils.rs:359-371 (each 16-byte half reduced mod p, then lo + hi * 2^64)."""
    lo = int.from_bytes(b32[0:16], "little") % P
    hi = int.from_bytes(b32[16:32], "little") % P
    return (lo + hi * POW2_64) % P
