"""Pure-Python BLAKE3 (hash mode). Host-side only: the reference derives every Poseidon constant
from blake3 (zk-lisp-proof-winterfell/src/poseidon/mod.rs:421-440, crate blake3 1.8.2); inputs are
a few dozen bytes, so speed is irrelevant. Checked against the official vectors in tests/."""
import struct

IV = (0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19)
PERM = (2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
CHUNK_START, CHUNK_END, PARENT, ROOT = 1, 2, 4, 8
M32 = 0xFFFFFFFF


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & M32


def _g(s, a, b, c, d, mx, my):
    s[a] = (s[a] + s[b] + mx) & M32
    s[d] = _rotr(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & M32
    s[b] = _rotr(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b] + my) & M32
    s[d] = _rotr(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & M32
    s[b] = _rotr(s[b] ^ s[c], 7)


def _compress(cv, block, counter, block_len, flags):
    m = list(struct.unpack("<16I", block))
    s = list(cv) + list(IV[:4]) + [counter & M32, (counter >> 32) & M32, block_len, flags]
    for r in range(7):
        _g(s, 0, 4, 8, 12, m[0], m[1]); _g(s, 1, 5, 9, 13, m[2], m[3])
        _g(s, 2, 6, 10, 14, m[4], m[5]); _g(s, 3, 7, 11, 15, m[6], m[7])
        _g(s, 0, 5, 10, 15, m[8], m[9]); _g(s, 1, 6, 11, 12, m[10], m[11])
        _g(s, 2, 7, 8, 13, m[12], m[13]); _g(s, 3, 4, 9, 14, m[14], m[15])
        if r < 6:
            m = [m[PERM[i]] for i in range(16)]
    return [s[i] ^ s[i + 8] for i in range(8)] + [s[i + 8] ^ cv[i] for i in range(8)]


def _chunk_cv(chunk: bytes, index: int, root: bool):
    cv = list(IV)
    nblocks = max(1, (len(chunk) + 63) // 64)
    out = None
    for bi in range(nblocks):
        blk = chunk[bi * 64:(bi + 1) * 64]
        flags = (CHUNK_START if bi == 0 else 0) | (CHUNK_END if bi == nblocks - 1 else 0)
        if root and bi == nblocks - 1:
            flags |= ROOT
        out = _compress(cv, blk.ljust(64, b"\0"), index, len(blk), flags)
        cv = out[:8]
    return out


def blake3(data: bytes) -> bytes:
    chunks = [data[i:i + 1024] for i in range(0, len(data), 1024)] or [b""]
    if len(chunks) == 1:
        words = _chunk_cv(chunks[0], 0, True)
        return struct.pack("<8I", *words[:8])
    stack = []
    for ci, ch in enumerate(chunks[:-1]):
        cur = _chunk_cv(ch, ci, False)[:8]
        total = ci + 1
        while total & 1 == 0:
            left = stack.pop()
            cur = _compress(IV, struct.pack("<16I", *(left + cur)), 0, 64, PARENT)[:8]
            total >>= 1
        stack.append(cur)
    cur = _chunk_cv(chunks[-1], len(chunks) - 1, False)[:8]
    words = None
    while stack:
        left = stack.pop()
        words = _compress(IV, struct.pack("<16I", *(left + cur)), 0, 64, PARENT | (0 if stack else ROOT))
        cur = words[:8]
    return struct.pack("<8I", *words[:8])
