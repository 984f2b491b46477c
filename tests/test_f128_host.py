"""CPU tests of the DEVICE limb arithmetic (zk-lisp_b200/csrc/f128.cuh, f128_lazy.cuh) compiled with g++: the PTX
carry-flag primitives are emulated on the host, so the exact limb algorithms the kernels run are checked against
Python integers — canonical ops, the lazy (non-canonical) ops of the tensor-core permutation and its epilogue."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from util import P, from_fe, rand_fe, to_fe

HERE = os.path.dirname(os.path.abspath(__file__))
M128 = (1 << 128) - 1


@pytest.fixture(scope="module")
def hostlib():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhost_f128.so")
    src = os.path.join(HERE, "host_f128.cc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def binop(lib, op, a, b):
    A, B = to_fe(a), to_fe(b)
    out = np.empty_like(A)
    lib.h_binop(C.c_int(op), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(len(a)))
    return from_fe(out)


def rand_u128(rng, n):
    edge = [0, 1, P - 1, P, P + 1, M128, M128 - 1, (1 << 128) - (45 << 40), (1 << 127), (1 << 64) - 1]
    return [rng.choice(edge) if rng.random() < 0.15 else rng.getrandbits(128) for _ in range(n)]


def test_canonical_ops(hostlib):
    rng = random.Random(11)
    a, b = rand_fe(rng, 5000), rand_fe(rng, 5000)
    assert binop(hostlib, 0, a, b) == [(x + y) % P for x, y in zip(a, b)]
    assert binop(hostlib, 1, a, b) == [(x - y) % P for x, y in zip(a, b)]
    assert binop(hostlib, 2, a, b) == [(x * y) % P for x, y in zip(a, b)]
    assert binop(hostlib, 4, a, b) == [pow(x, 3, P) for x in a]
    assert binop(hostlib, 5, a, b) == [(x + (y << 64)) % P for x, y in zip(a, b)]
    u = rand_u128(rng, 2000)
    assert binop(hostlib, 6, u, u) == [x % P for x in u]


def test_lazy_ops_accept_any_128_bit_value(hostlib):
    rng = random.Random(12)
    a, b = rand_u128(rng, 20000), rand_u128(rng, 20000)
    got = binop(hostlib, 7, a, b)
    assert all(g <= M128 for g in got)
    assert [g % P for g in got] == [(x * y) % P for x, y in zip(a, b)]
    got = binop(hostlib, 11, a, b)          # multiply-accumulate form of the first fold
    assert [g % P for g in got] == [(x * y) % P for x, y in zip(a, b)]
    got = binop(hostlib, 8, a, b)
    assert [g % P for g in got] == [(x + y) % P for x, y in zip(a, b)]
    assert binop(hostlib, 9, a, a) == [x % P for x in a]
    # worst cases for the carry fix-ups: everything at the top of the range
    top = [M128 - k for k in range(64)]
    got = binop(hostlib, 7, top, top[::-1])
    assert [g % P for g in got] == [(x * y) % P for x, y in zip(top, top[::-1])]
    got = binop(hostlib, 11, top, top[::-1])
    assert [g % P for g in got] == [(x * y) % P for x, y in zip(top, top[::-1])]
    got = binop(hostlib, 8, top, top[::-1])
    assert [g % P for g in got] == [(x + y) % P for x, y in zip(top, top[::-1])]


def test_mul_wide_is_the_exact_256_bit_product(hostlib):
    rng = random.Random(13)
    a, b = rand_u128(rng, 4000), rand_u128(rng, 4000)
    A, B = to_fe(a), to_fe(b)
    out = np.empty((len(a), 8), dtype=np.uint32)
    hostlib.h_mul_wide(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(len(a)))
    got = [sum(int(out[i, j]) << (32 * j) for j in range(8)) for i in range(len(a))]
    assert got == [x * y for x, y in zip(a, b)]


def test_sqr_wide_matches_product(hostlib):
    rng = random.Random(16)
    a = rand_u128(rng, 8000) + [M128 - k for k in range(32)] + [((1 << 32) - 1) << (32 * k) for k in range(4)]
    A = to_fe(a)
    out = np.empty((len(a), 8), dtype=np.uint32)
    hostlib.h_sqr_wide(A.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(len(a)))
    got = [sum(int(out[i, j]) << (32 * j) for j in range(8)) for i in range(len(a))]
    assert got == [x * x for x in a]
    assert [g % P for g in binop(hostlib, 10, a, a)] == [x * x % P for x in a]


def test_tc_epilogue_folds_byte_accumulators(hostlib):
    """acc[u] <= TC_ACC_MAX (pairs fit 32 bits; checked against the loaded constants by zkl_suite_load);
    result ≡ sum acc[u] 2^(8u) + rc (mod p)."""
    rng = random.Random(14)
    hostlib.h_tc_acc_max.restype = C.c_uint32
    n, amax = 20000, int(hostlib.h_tc_acc_max())
    assert (amax << 8) + amax < 2**32
    acc = np.empty((n, 16), dtype=np.uint32)
    for i in range(n):
        mode = rng.random()
        for u in range(16):
            acc[i, u] = amax if mode < 0.05 else (0 if mode < 0.1 else rng.randrange(amax + 1))
    rc = rand_fe(rng, n)
    RC = to_fe(rc)
    out = np.empty_like(RC)
    hostlib.h_tc_epilogue(acc.ctypes.data_as(C.c_void_p), RC.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(n))
    got = from_fe(out)
    want = [(sum(int(acc[i, u]) << (8 * u) for u in range(16)) + rc[i]) % P for i in range(n)]
    assert [g % P for g in got] == want
    hostlib.h_tc_epilogue(acc.ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p), C.c_long(n))   # no addend (NTT stages)
    assert [g % P for g in from_fe(out)] == [sum(int(acc[i, u]) << (8 * u) for u in range(16)) % P for i in range(n)]


def test_mds_byte_decomposition_identity():
    """The identity behind tcperm.cuh in Python integers: MDS*c == sum_u 2^(8u) sum_{k,v} byte_u(M[i][k] 2^(8v) mod p) byte_v(c_k)."""
    rng = random.Random(15)
    mds = [[rng.randrange(P) for _ in range(12)] for _ in range(12)]
    cubes = [rng.getrandbits(256) for _ in range(12)]          # raw 256-bit cubes (lazy: not reduced)
    for i in range(12):
        acc = [0] * 16
        for k in range(12):
            for v in range(32):
                t = (mds[i][k] << (8 * v)) % P
                cb = (cubes[k] >> (8 * v)) & 0xFF
                for u in range(16):
                    acc[u] += ((t >> (8 * u)) & 0xFF) * cb
        assert max(acc) < 2**31
        assert sum(a << (8 * u) for u, a in enumerate(acc)) % P == sum(mds[i][k] * cubes[k] for k in range(12)) % P
