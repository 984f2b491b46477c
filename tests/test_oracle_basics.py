"""CPU tests: pin the oracle's arithmetic, BLAKE3, parameter derivation, hasher, FFT and Merkle tree
against independent Python big-int restatements and the BLAKE3 official vectors."""
import json
import os
import random

import numpy as np

import pyref
from util import P, from_fe, rand_fe, to_fe

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_field_ops_match_python_ints(oracle):
    rng = random.Random(7)
    a, b = rand_fe(rng, 4000), rand_fe(rng, 4000)
    A, B = to_fe(a), to_fe(b)
    assert from_fe(oracle.f_binop(0, A, B)) == [(x + y) % P for x, y in zip(a, b)]
    assert from_fe(oracle.f_binop(1, A, B)) == [(x - y) % P for x, y in zip(a, b)]
    assert from_fe(oracle.f_binop(2, A, B)) == [(x * y) % P for x, y in zip(a, b)]
    assert from_fe(oracle.f_binop(4, A, B)) == [pow(x, 3, P) for x in a]
    assert from_fe(oracle.f_binop(3, A[:200], B[:200])) == [pow(x, P - 2, P) for x in a[:200]]


def test_field_constants(zkl):
    f = zkl.field
    assert f.P == 340282366920938463463374557953744961537
    # winter-math f128 TWO_ADIC_ROOT_OF_UNITY
    assert f.TWO_ADIC_ROOT == 23953097886125630542083529559205016746
    assert pow(f.TWO_ADIC_ROOT, 1 << 39, f.P) == f.P - 1
    assert f.root_of_unity(1) == f.P - 1


def test_blake3_official_vectors(oracle, zkl):
    kat = json.load(open(os.path.join(GOLDEN, "blake3_kat.json")))
    for case in kat["cases"]:
        data = bytes(i % 251 for i in range(case["input_len"]))
        assert oracle.blake3(data).hex() == case["hash"]
        assert zkl.blake3.blake3(data).hex() == case["hash"]
    assert oracle.blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"
    assert oracle.blake3(b"abc").hex() == "6437b3ac38465133ffb63b75273a8db548c558465d79db03fd359c6cd5bd9d85"


def test_suite_derivation_c_matches_python(oracle, zkl):
    import ctypes as C
    for sid in (bytes(32), bytes([42] * 32), bytes(range(32))):
        s = zkl.capi.Suite()
        oracle.load().zo_derive_suite(sid, C.byref(s))
        py = zkl.params.get_poseidon_suite(sid)
        assert [s.dom[i].int() for i in range(2)] == py.dom
        assert [[s.mds[i][j].int() for j in range(12)] for i in range(12)] == py.mds
        assert [[s.rc[r][j].int() for j in range(12)] for r in range(27)] == py.rc
        # Cauchy MDS: entries are inverses of x_i + y_j, all non-zero (poseidon/mod.rs:450-457)
        assert all(v != 0 for row in py.mds for v in row)
        r = zkl.capi.RomSuite()
        oracle.load().zo_derive_rom_suite(sid, C.byref(r))
        mds, rc = zkl.params.get_rom_suite(sid)
        assert [[r.mds[i][j].int() for j in range(3)] for i in range(3)] == mds
        assert [[r.rc[k][j].int() for j in range(3)] for k in range(27)] == rc


def test_hasher_matches_independent_python(oracle, zkl):
    rng = random.Random(3)
    # permutation
    st = rand_fe(rng, 24)
    out = from_fe(oracle.permute(to_fe(st).reshape(2, 12, 2)))
    suite = zkl.params.get_poseidon_suite(bytes(32))
    assert out == zkl.params.permute(suite, st[:12]) + zkl.params.permute(suite, st[12:])
    # hash_elements: empty, odd tail, exactly 10 absorbs (19 elems -> 1 + 10 = 11), 9 -> 1+5
    for ln in (0, 1, 2, 7, 17, 18, 19, 20, 39, 40):
        e = rand_fe(rng, ln)
        got = oracle.hash_elements(to_fe(e).reshape(1, ln, 2))[0].tobytes()
        assert got == pyref.hash_elements(e), ln
    d = [pyref.digest(x) for x in rand_fe(rng, 16)]
    assert oracle.merge(np.frombuffer(d[0] + d[1], dtype=np.uint8))[0].tobytes() == pyref.merge(d[0], d[1])
    for v in (0, 1, 2**64 - 1, 12345):
        assert oracle.merge_with_int(d[2], [v])[0].tobytes() == pyref.merge_with_int(d[2], v)
    for cnt in (1, 2, 9, 10, 14, 16):
        assert oracle.merge_many(np.frombuffer(b"".join(d[:cnt]), dtype=np.uint8)) == pyref.merge_many(d[:cnt])
    for data in (b"", b"x", bytes(range(31)), bytes(range(32)), bytes(range(200))):
        assert oracle.hash_bytes(data) == pyref.hash_bytes(data)
    # digests carry 16 significant bytes (utils.rs:375-381)
    assert pyref.hash_bytes(b"abc")[16:] == bytes(16)


def test_fft_roundtrip_and_definition(oracle, zkl):
    rng = random.Random(5)
    n, blowup = 16, 4
    coeffs = rand_fe(rng, n)
    ev = from_fe(oracle.evaluate_columns(to_fe(coeffs).reshape(1, n, 2), blowup))
    w = zkl.field.root_of_unity(6)
    for i in range(n * blowup):
        x = 3 * pow(w, i, P) % P
        assert ev[i] == sum(c * pow(x, j, P) for j, c in enumerate(coeffs)) % P
    # interpolate is the inverse of evaluation over <w_n>
    wn = zkl.field.root_of_unity(4)
    vals = [sum(c * pow(wn, i * j, P) for j, c in enumerate(coeffs)) % P for i in range(n)]
    assert from_fe(oracle.interpolate_columns(to_fe(vals).reshape(1, n, 2))) == coeffs


def test_merkle_tree_matches_path_fold(oracle):
    rng = random.Random(9)
    leaves = [pyref.digest(x) for x in rand_fe(rng, 8)]
    root, nodes = oracle.merkle_build(np.frombuffer(b"".join(leaves), dtype=np.uint8))
    lvl = leaves
    while len(lvl) > 1:
        lvl = [pyref.merge(lvl[2 * i], lvl[2 * i + 1]) for i in range(len(lvl) // 2)]
    assert root == lvl[0]
    assert nodes[1].tobytes() == root
    assert nodes[4].tobytes() == pyref.merge(leaves[0], leaves[1])
