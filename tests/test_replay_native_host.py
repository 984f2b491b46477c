"""CPU tests of the NATIVE replay code's host logic (zk-lisp_b200/csrc/api_replay.inl): the proof parser and the merge-plan builder,
compiled with g++ through tests/host_replay.cc, must agree with the Python restatement (step.InnerProof.parse, replay.MergePlan)
on real proofs — same spans, same schedule index for index. The device half of zkl_replay_children runs in the GPU suite."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_child_replay import _step

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hr():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "host_replay.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", os.path.join(HERE, "host_replay.cc"), "-o", so])
    return C.CDLL(so)


def _parse(hr, proof: bytes):
    buf = np.frombuffer(proof, dtype=np.uint8)
    fields, nonce, spans = np.zeros(14, dtype=np.uint64), C.c_uint64(), np.zeros(2 * 100, dtype=np.uint64)
    rc = hr.hr_parse(C.c_void_p(buf.ctypes.data), C.c_size_t(len(proof)), C.c_void_p(fields.ctypes.data), C.byref(nonce),
                     C.c_void_p(spans.ctypes.data), C.c_size_t(100))
    return rc, [int(x) for x in fields], nonce.value, [(int(spans[2 * k]), int(spans[2 * k + 1])) for k in range(100)]


def test_native_parser_matches_python_view(zkl, oracle, hr):
    for n, seed, kw in ((64, 7, {}), (32, 9, dict(queries=40, parts=(4, 8), grind=0))):
        step = _step(zkl, oracle, n, seed, **kw)[0]
        proof, v = step.proof.inner_bytes, step.proof.inner
        rc, f, nonce, spans = _parse(hr, proof)
        assert rc == 0 and nonce == v.pow_nonce
        assert f == [v.main_width, v.aux_width, v.trace_len.bit_length() - 1, v.num_constraints, v.num_queries, v.blowup, v.grinding,
                     v.field_extension, v.fri_folding, v.fri_remainder_max_degree, v.num_partitions, v.hash_rate, v.num_unique_queries,
                     len(v.fri_layers)]
        expect = [v.commitments, v.trace_query_values, v.trace_query_paths, v.constraint_query_values, v.constraint_query_paths,
                  v.ood_trace, v.ood_constraints, v.fri_remainder] + [x for layer in v.fri_layers for x in layer]
        assert [proof[o:o + ln] for o, ln in spans[:len(expect)]] == expect
        # every truncation and a trailing byte are refused by both parsers
        for cut in list(range(0, len(proof), 997)) + [len(proof) - 1]:
            assert _parse(hr, proof[:cut])[0] == 1
            with pytest.raises(zkl.step.InvalidInput):
                zkl.step.InnerProof.parse(proof[:cut])
        assert _parse(hr, proof + b"\x00")[0] == 1


def _native_plan(hr, openings):
    """openings: [(paths bytes, leaves [32-byte], positions)] -> flattened schedule or the index of the malformed opening."""
    num = len(openings)
    paths = [np.frombuffer(p, dtype=np.uint8) for p, _, _ in openings]
    leaves = [np.frombuffer(b"".join(l), dtype=np.uint8) for _, l, _ in openings]
    pos = [np.array(p, dtype=np.uint32) for _, _, p in openings]
    arr = lambda xs: (C.c_void_p * num)(*[x.ctypes.data for x in xs])
    cap = 1 << 16
    known, nk = np.zeros((cap, 32), dtype=np.uint8), C.c_size_t()
    left, right = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
    counts, want = np.zeros(64, dtype=np.uint32), np.zeros(num, dtype=np.uint32)
    rc = hr.hr_plan(C.c_size_t(num), arr(paths), (C.c_size_t * num)(*[len(p) for p, _, _ in openings]), arr(leaves), arr(pos),
                    (C.c_size_t * num)(*[len(p) for _, _, p in openings]), C.c_void_p(known.ctypes.data), C.byref(nk),
                    C.c_void_p(left.ctypes.data), C.c_void_p(right.ctypes.data), C.c_void_p(counts.ctypes.data), C.c_void_p(want.ctypes.data))
    if rc < 0:
        return -1 - rc
    total = int(counts[:rc].sum())
    return known[:nk.value], left[:total], right[:total], counts[:rc], want


def test_native_plan_builder_matches_python_plan(zkl, oracle, hr):
    R = zkl.replay
    kids = [_step(zkl, oracle, 64, 7)[0], _step(zkl, oracle, 32, 9, queries=40, parts=(4, 8), grind=0)[0]]
    openings, plan = [], R.MergePlan()
    for step in kids:
        v = step.proof.inner
        fs = R.replay_fs_from_step(oracle, step)
        k = len(fs.query_positions)
        rows = np.frombuffer(v.trace_query_values, dtype=np.uint64).reshape(k, v.main_width, 2)
        leaves = R.hash_rows(oracle, rows, R.partition_size(v.num_partitions, v.hash_rate, v.main_width))
        openings.append((v.trace_query_paths, leaves, fs.query_positions))
        positions, domain = list(fs.query_positions), v.lde_domain_size
        for values, paths in v.fri_layers:
            positions = R.fold_positions(positions, domain)
            pairs = np.frombuffer(values, dtype=np.uint64).reshape(len(positions), 2, 2)
            openings.append((paths, [d.tobytes() for d in oracle.hash_elements(pairs)], positions))
            domain //= 2
    for paths, leaves, positions in openings:
        plan.add_opening(*R.read_batch_proof(paths), leaves, positions, bytes(32), "x")
    got, expect = _native_plan(hr, openings), plan.flatten()
    assert len(got) == len(expect) == 5
    for g, e in zip(got, expect):
        assert np.array_equal(g, e)
    # the schedule authenticates: running it gives each tree's committed root
    from test_child_replay import _PlanHasher
    roots = _PlanHasher(oracle).merge_plan(*got)
    v0 = kids[0].proof.inner
    assert roots[0].tobytes() == v0.parse_commitments()[0][0] and roots[1].tobytes() == v0.parse_commitments()[2][0]
    # malformed openings are refused and identified: a dropped node, a position outside the tree, a missing leaf
    paths, leaves, positions = openings[0]
    assert _native_plan(hr, [openings[1], (paths[:-32], leaves, positions)]) == 1
    assert _native_plan(hr, [(paths, leaves, positions[:-1] + [1 << 20])]) == 0
    assert _native_plan(hr, [(paths, leaves[:-1], positions[:-1])]) == 0
    for bad in ((paths[:-32], leaves, positions), (paths, leaves[:-1], positions[:-1])):
        with pytest.raises(zkl.step.InvalidInput):
            R.MergePlan().add_opening(*R.read_batch_proof(bad[0]), bad[1], bad[2], bytes(32), "x")
