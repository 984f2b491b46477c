"""ctypes loader for the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    global _lib
    if _lib is None:
        build()   # no-op when up to date; rebuilds when the library came from a host with another CPU (oracle/Makefile)
        _lib = C.CDLL(LIB)
        _lib.zo_prove_many.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f_binop(op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    load().zo_f_binop(op, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def blake3(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    load().zo_blake3(data, C.c_size_t(len(data)), out)
    return out.raw


def permute(states: np.ndarray) -> np.ndarray:
    st = np.ascontiguousarray(states, dtype=np.uint64).copy()
    load().zo_permute(_p(st), C.c_size_t(st.shape[0]))
    return st


def hash_elements(elems: np.ndarray) -> np.ndarray:
    e = np.ascontiguousarray(elems, dtype=np.uint64)
    out = np.zeros((e.shape[0], 32), dtype=np.uint8)
    load().zo_hash_elements(_p(e), C.c_size_t(e.shape[1]), C.c_size_t(e.shape[0]), _p(out))
    return out


def merge(pairs: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(pairs, dtype=np.uint8).reshape(-1, 64)
    out = np.zeros((p.shape[0], 32), dtype=np.uint8)
    load().zo_merge(_p(p), C.c_size_t(p.shape[0]), _p(out))
    return out


def merge_with_int(seed: bytes, values) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.uint64)
    out = np.zeros((v.shape[0], 32), dtype=np.uint8)
    load().zo_merge_with_int(bytes(seed), _p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def hash_bytes(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    load().zo_hash_bytes(bytes(data), C.c_size_t(len(data)), out)
    return out.raw


def merge_many(digests: np.ndarray) -> bytes:
    d = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
    out = C.create_string_buffer(32)
    load().zo_merge_many(_p(d), C.c_size_t(d.shape[0]), out)
    return out.raw


def merkle_build(leaves: np.ndarray):
    lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    root = C.create_string_buffer(32)
    nodes = np.zeros((lv.shape[0], 32), dtype=np.uint8)
    load().zo_merkle_build(_p(lv), C.c_size_t(lv.shape[0]), root, _p(nodes))
    return root.raw, nodes


def interpolate_columns(cols: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(cols, dtype=np.uint64).copy()
    load().zo_interpolate_columns(_p(c), C.c_size_t(c.shape[0]), C.c_size_t(c.shape[1]))
    return c


def evaluate_columns(polys: np.ndarray, blowup: int, offset: int = 3) -> np.ndarray:
    p = np.ascontiguousarray(polys, dtype=np.uint64)
    out = np.zeros((p.shape[0], p.shape[1] * blowup, 2), dtype=np.uint64)
    off = np.array([offset & (2**64 - 1), offset >> 64], dtype=np.uint64)
    load().zo_evaluate_columns(_p(p), C.c_size_t(p.shape[0]), C.c_size_t(p.shape[1]), C.c_size_t(blowup), _p(off), _p(out))
    return out


def trace_lde_commit(trace: np.ndarray, opts_c, want_lde=False):
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    W, n = t.shape[0], t.shape[1]
    root = C.create_string_buffer(32)
    polys = np.zeros((W, n, 2), dtype=np.uint64)
    lde = np.zeros((W, n * opts_c.blowup, 2), dtype=np.uint64) if want_lde else None
    load().zo_trace_lde_commit(_p(t), C.c_size_t(W), C.c_size_t(n), C.byref(opts_c), root, _p(polys),
                               _p(lde) if want_lde else None)
    return root.raw, polys, lde


def eval_constraints(trace: np.ndarray, opts_c, air_c, coeffs: np.ndarray) -> np.ndarray:
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    W, n = t.shape[0], t.shape[1]
    cc = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.zeros((n * air_c.ce_blowup, 2), dtype=np.uint64)
    load().zo_eval_constraints(_p(t), C.c_size_t(W), C.c_size_t(n), C.byref(opts_c), C.byref(air_c), _p(cc), _p(out))
    return out


def eval_transition_frame(air_c, cur, nxt, periodic) -> np.ndarray:
    cur = np.ascontiguousarray(cur, dtype=np.uint64); nxt = np.ascontiguousarray(nxt, dtype=np.uint64)
    per = np.ascontiguousarray(periodic, dtype=np.uint64)
    out = np.zeros((air_c.num_transition, 2), dtype=np.uint64)
    load().zo_eval_transition_frame(C.byref(air_c), _p(cur), _p(nxt), _p(per), _p(out))
    return out


def periodic_at(n: int, x: int) -> np.ndarray:
    xa = np.array([x & (2**64 - 1), x >> 64], dtype=np.uint64)
    out = np.zeros((32, 2), dtype=np.uint64)
    load().zo_periodic_at(C.c_size_t(n), _p(xa), _p(out))
    return out


def constraint_commit(ce_evals: np.ndarray, n, ce_blowup, num_cols, opts_c):
    e = np.ascontiguousarray(ce_evals, dtype=np.uint64)
    root = C.create_string_buffer(32)
    poly = np.zeros((num_cols, n, 2), dtype=np.uint64)
    load().zo_constraint_commit(_p(e), C.c_size_t(n), C.c_size_t(ce_blowup), C.c_size_t(num_cols), C.byref(opts_c), root, _p(poly))
    return root.raw, poly


def prove_segment(trace: np.ndarray, air_c, seed: np.ndarray, opts_c, summary=None):
    t = np.ascontiguousarray(trace, dtype=np.uint64)
    W, n = t.shape[0], t.shape[1]
    s = np.ascontiguousarray(seed, dtype=np.uint64)
    cap = C.c_size_t(4 << 20)
    buf = np.zeros(cap.value, dtype=np.uint8)
    rc = load().zo_prove_segment(_p(t), C.c_size_t(W), C.c_size_t(n), C.byref(air_c), _p(s), C.c_size_t(s.shape[0]),
                                 C.byref(opts_c), _p(buf), C.byref(cap), C.byref(summary) if summary is not None else None)
    assert rc == 0
    return buf[:cap.value].tobytes()


def verify_segment(proof: bytes, W, n, air_c, seed: np.ndarray, opts_c):
    s = np.ascontiguousarray(seed, dtype=np.uint64)
    msg = C.create_string_buffer(256)
    rc = load().zo_verify_segment(proof, C.c_size_t(len(proof)), C.c_size_t(W), C.c_size_t(n), C.byref(air_c), _p(s),
                                  C.c_size_t(s.shape[0]), C.byref(opts_c), msg, C.c_size_t(256))
    return rc, msg.value.decode()


def prove_many(traces: np.ndarray, air_c, seed: np.ndarray, opts_c, threads: int) -> float:
    t = np.ascontiguousarray(traces, dtype=np.uint64)
    count, W, n = t.shape[0], t.shape[1], t.shape[2]
    s = np.ascontiguousarray(seed, dtype=np.uint64)
    return load().zo_prove_many(_p(t), C.c_size_t(count), C.c_size_t(W), C.c_size_t(n), C.byref(air_c), _p(s),
                                C.c_size_t(s.shape[0]), C.byref(opts_c), C.c_int(threads))


def batch_root(bp: dict, leaves: np.ndarray, positions) -> bytes:
    """Root recomputed from a batch Merkle opening (depth, node lists) for `leaves` at `positions`; b"" if malformed."""
    lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    pos = np.ascontiguousarray(positions, dtype=np.uint32)
    lens = np.array([len(l) for l in bp["nodes"]], dtype=np.uint32)
    flat = np.frombuffer(b"".join(b"".join(l) for l in bp["nodes"]) or bytes(32), dtype=np.uint8).copy()
    out = C.create_string_buffer(32)
    rc = load().zo_batch_root(C.c_int(bp["depth"]), C.c_size_t(len(bp["nodes"])), _p(lens), _p(flat), _p(lv), _p(pos),
                              C.c_size_t(pos.shape[0]), out)
    return out.raw if rc == 0 else b""
