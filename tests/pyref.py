"""Pure-Python (big-int) restatement of the hasher, written independently of oracle/ from
zk-lisp-proof-winterfell/src/poseidon/hasher.rs:144-231. Small cases only; used to pin the C oracle."""
from zk_lisp_b200 import params
from zk_lisp_b200.field import P, fold_bytes32_to_fe


def sponge(domain: str, data: bytes, suite=None) -> int:
    ps = suite or params.get_poseidon_suite(bytes(32))
    st = [0] * 12
    st[10], st[11] = ps.dom
    lane = 0

    def absorb(m):
        nonlocal st, lane
        st[lane] = (st[lane] + m) % P
        lane += 1
        if lane == 10:
            st = params.permute(ps, st)
            lane = 0

    absorb(fold_bytes32_to_fe(domain.encode()[:32].ljust(32, b"\0")))
    for i in range(0, len(data), 32):
        absorb(fold_bytes32_to_fe(data[i:i + 32].ljust(32, b"\0")))
    if lane != 0:
        st = params.permute(ps, st)
    return st[0]


def digest(fe: int) -> bytes:
    return fe.to_bytes(16, "little") + bytes(16)


def hash_elements(elems) -> bytes:
    return digest(sponge("winter/hash/elements", b"".join(e.to_bytes(16, "little") for e in elems)))


def merge(a: bytes, b: bytes) -> bytes:
    return digest(sponge("zkl/winter/hash/merge", a + b))


def merge_many(ds) -> bytes:
    if not ds:
        return bytes(32)
    return digest(sponge("zkl/winter/hash/merge_many", b"".join(ds)))


def merge_with_int(seed: bytes, v: int) -> bytes:
    return digest(sponge("zkl/winter/hash/merge_with_int", seed + v.to_bytes(8, "little")))


def hash_bytes(data: bytes) -> bytes:
    return digest(sponge("zkl/winter/hash/bytes", data))
