"""Deterministic derivation of the Poseidon-style parameters, host side.

Mirrors /root/reference/zk-lisp-proof-winterfell/src/poseidon/mod.rs (suite derivation),
src/utils.rs:114-141 (ROM weights) and src/commit.rs:31-79 (program_field_commitment).
The Rust side of a real integration would hand these constants to `zkl_suite_load`; this module
is the stand-in so that the Python host layer is self-contained. Parameter derivation runs once
per program on the host in the reference too — it is not part of the proving hot path.
"""
from functools import lru_cache

from .blake3 import blake3
from .field import P, POW2_64, inv

POSEIDON_ROUNDS = 27  # vm/layout.rs:22
HASHER_SUITE_ID = bytes(32)  # poseidon/hasher.rs:23


def ro_from_slices(domain: str, parts) -> int:
    """poseidon/mod.rs:421-440."""
    d = blake3(domain.encode() + b"".join(parts))
    lo = int.from_bytes(d[0:8], "little")
    hi = int.from_bytes(d[8:16], "little")
    return (lo + hi * POW2_64) % P


def _derive_points(domain: str, sid: bytes, n: int):
    """poseidon/mod.rs:114-131."""
    pts, ctr = [], 0
    while len(pts) < n:
        cand = ro_from_slices(domain, [sid, bytes([len(pts)]), ctr.to_bytes(4, "little")])
        if cand != 0 and cand not in pts:
            pts.append(cand)
        else:
            ctr = (ctr + 1) & 0xFFFFFFFF
    return pts


class Suite:
    """poseidon/mod.rs:33-38 PoseidonSuite (t = 12, rate 10)."""

    def __init__(self, dom, mds, rc):
        self.dom, self.mds, self.rc = dom, mds, rc


@lru_cache(maxsize=None)
def get_poseidon_suite(suite_id: bytes) -> Suite:
    """poseidon/mod.rs:52-75, 111-217."""
    dom = [ro_from_slices("zkl/poseidon2/dom/c0", [suite_id]), ro_from_slices("zkl/poseidon2/dom/c1", [suite_id])]
    x = _derive_points("zkl/poseidon2/mds/x", suite_id, 12)
    y = _derive_points("zkl/poseidon2/mds/y", suite_id, 12)
    adj = 0
    while any((xi + yj) % P == 0 for xi in x for yj in y):
        for j in range(12):
            cand = ro_from_slices("zkl/poseidon2/mds/y", [suite_id, bytes([j]), adj.to_bytes(4, "little")])
            y[j] = 1 if cand == 0 else cand
        adj += 1
        if adj > 1_000_000:
            raise RuntimeError("poseidon: failed to derive valid 12x12 MDS")
    mds = [[inv((x[i] + y[j]) % P) for j in range(12)] for i in range(12)]
    rc = [[ro_from_slices("zkl/poseidon2/rc", [suite_id, bytes([r]), bytes([lane])]) for lane in range(12)]
          for r in range(POSEIDON_ROUNDS)]
    return Suite(dom, mds, rc)


@lru_cache(maxsize=None)
def get_rom_suite(suite_id: bytes):
    """poseidon/mod.rs:186-197, 219-251 -> (mds 3x3, rc 27x3)."""
    x = _derive_points("zkl/rom3/mds/x", suite_id, 3)
    y = _derive_points("zkl/rom3/mds/y", suite_id, 3)
    mds = [[inv((x[i] + y[j]) % P) for j in range(3)] for i in range(3)]
    rc = [[ro_from_slices("zkl/rom3/rc", [suite_id, bytes([r]), bytes([lane])]) for lane in range(3)]
          for r in range(POSEIDON_ROUNDS)]
    return mds, rc


ROM_W_SEED_0, ROM_W_SEED_1 = 17, 1037  # utils.rs:27-28


def rom_weights_for_seed(seed: int):
    """utils.rs:114-141: [3^(seed+1) .. 3^(seed+59)]."""
    return [pow(3, seed + 1 + k, P) for k in range(59)]


def permute(suite: Suite, state):
    """poseidon/hasher.rs:173-190 on Python ints (parameter derivation / tiny host checks only)."""
    s = list(state)
    for rc in suite.rc:
        c = [pow(v, 3, P) for v in s]
        s = [(sum(suite.mds[i][k] * c[k] for k in range(12)) + rc[i]) % P for i in range(12)]
    return s


def program_field_commitment(blake32: bytes):
    """commit.rs:31-79."""
    suite = get_poseidon_suite(bytes(blake32))

    def fe16(b):
        return (int.from_bytes(b[0:8], "little") + int.from_bytes(b[8:16], "little") * POW2_64) % P

    s = [0] * 12
    s[0], s[1] = fe16(blake32[0:16]), fe16(blake32[16:32])
    s[10], s[11] = suite.dom
    s = permute(suite, s)
    return [s[0], s[1]]
