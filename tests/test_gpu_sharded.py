"""Column-sharded proving on the GPU (zk-lisp_b200/sharded.py, include/zkl_b200.h: zkl_comm).

* the NCCL commitment flow with a world of ONE rank gives the root of the single-GPU hook and of the CPU oracle;
* the column-sharded PROOF of one segment, with 1, 2, 4 and 8 ranks as threads of this process on one GPU (`zkl_comm_local_*`: the
  library's whole multi-rank flow — balanced partition ownership, partition-digest all-to-all, per-rank sub-trees, CE-coset
  all-to-all, DEEP reduce-scatter, FRI per coset range, one-contributor openings — with device-to-device copies as the transport),
  must be byte-identical on every rank to the single-GPU proof, which the other tests pin against the CPU checker.
The same flow over NCCL (one process per GPU) runs under torchrun: tools/sharded_prove_check.py; its index logic under gloo on CPU."""
import dataclasses
import socket

import numpy as np
import pytest

import synth
from conftest import HASH_MODES
from util import fast_random_trace

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nccl_world1():
    import torch
    import torch.distributed as dist
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist
    dist.destroy_process_group()


@pytest.mark.parametrize("width,n,parts", [(40, 64, (4, 8)), (219, 128, (16, 16)), (33, 256, (3, 4))])
def test_sharded_commit_matches_single_gpu_and_oracle(ctx, zkl, oracle, nccl_world1, width, n, parts):
    if ctx.hash_mode == "int-pipe":
        pytest.skip("same code path as tc-forced above the ABI")
    trace = fast_random_trace(width, n, seed=width + n)
    opts = zkl.air.ProofOptions(num_queries=8, blowup=16, grinding=0, num_partitions=parts[0], hash_rate=parts[1])
    sc = zkl.sharded.ShardedTraceCommit(ctx, nccl_world1, opts)
    root, plan = sc.commit(trace, width, n)
    assert plan.num_partitions >= 2 and plan.columns(0) == (0, width)
    oroot, _, _ = oracle.trace_lde_commit(trace, opts.to_c(), want_lde=False)
    assert root == oroot
    prover = zkl.ZkProver(ctx, opts, None)
    lde, _ = prover.new_trace_lde(trace)
    assert lde.get_main_trace_commitment() == root
    lde.free()


def _rank_contexts(zkl, world, mode):
    thr, path, coop = HASH_MODES[mode]
    out = []
    for _ in range(world):
        c = zkl.Context(0)
        if thr is not None:
            c._check(c.lib.zkl_set_bulk_threshold(c.h, thr))
        c._check(c.lib.zkl_set_hash_path(c.h, path))
        c._check(c.lib.zkl_set_coop_path(c.h, coop))
        out.append(c)
    return out


CASES = {
    "vm204": (lambda: synth.valid_segment(n=256, queries=16, blowup=16, grind=6, seed=5), (16, 16)),             # 13 partitions of 16
    "full219": (lambda: synth.rich_segment(n=1024, seed=23, merkle=True, queries=16, grind=6), (16, 16)),        # X's shape: 14 partitions
    "narrow": (lambda: synth.valid_segment(n=128, queries=8, blowup=8, grind=4, seed=9), (32, 8)),               # blowup 8, 8-column partitions
}


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("case", list(CASES))
def test_sharded_proof_is_the_single_gpu_proof(ctx, zkl, case, world):
    """Every rank's bytes == zkl_prove_segment_pi on the whole trace, with every tree and FRI layer exchanged (min_leaves = 0) and
    with the production threshold (small trees built on every rank)."""
    gen, parts = CASES[case]
    trace, air_pi, opts = gen()
    opts = dataclasses.replace(opts, num_partitions=parts[0], hash_rate=parts[1])
    single = zkl.ZkProver(ctx, opts, air_pi).prove(trace)
    ranks = _rank_contexts(zkl, world, ctx.hash_mode)
    try:
        for min_leaves in (0, 4096):
            proofs = zkl.sharded.prove_segment_thread_ranks(ranks, trace, opts, air_pi, min_leaves=min_leaves)
            for r, p in enumerate(proofs):
                assert p == single, f"rank {r} of {world}, min_leaves {min_leaves}: proof differs from the single-GPU proof"
    finally:
        for c in ranks:
            c.close()


def test_sharded_proof_two_pass_sizes(ctx, zkl):
    """n = 2^17: the transforms run as two passes over HBM and the boundary polynomials are extended over the rank's CE cosets only
    (dev_lde_cosets); millions of composition coefficients are drawn by the bulk coin kernel. 4 ranks, production thresholds."""
    if ctx.hash_mode != "default":
        pytest.skip("one hashing mode is enough at this size")
    trace, air_pi, opts = synth.valid_segment(n=1 << 17, queries=32, blowup=16, grind=8, seed=77, full_layout=True, poseidon_levels=False)
    opts = dataclasses.replace(opts, num_partitions=16, hash_rate=16)
    single = zkl.ZkProver(ctx, opts, air_pi).prove(trace)
    ranks = _rank_contexts(zkl, 4, ctx.hash_mode)
    try:
        proofs = zkl.sharded.prove_segment_thread_ranks(ranks, trace, opts, air_pi)
        assert all(p == single for p in proofs)
    finally:
        for c in ranks:
            c.close()


def test_sharded_refuses_what_it_cannot_shard(ctx, zkl):
    import ctypes as C
    trace, air_pi, opts = synth.valid_segment(n=64, queries=8, blowup=16, grind=4)
    lib = ctx.lib
    f, w = C.c_size_t(), C.c_size_t()
    one = dataclasses.replace(opts, num_partitions=1, hash_rate=0).to_c()          # whole rows are hashed: nothing to shard by
    assert lib.zkl_shard_columns(C.byref(one), trace.shape[0], 0, 2, C.byref(f), C.byref(w)) != 0
    many = dataclasses.replace(opts, num_partitions=16, hash_rate=16).to_c()       # 13 partitions
    assert lib.zkl_shard_columns(C.byref(many), trace.shape[0], 0, 16, C.byref(f), C.byref(w)) != 0   # more ranks than partitions
    cover = []
    for r in range(8):
        assert lib.zkl_shard_columns(C.byref(many), trace.shape[0], r, 8, C.byref(f), C.byref(w)) == 0
        assert f.value % 16 == 0 and w.value > 0
        cover += list(range(f.value, f.value + w.value))
    assert cover == list(range(trace.shape[0]))
    # a rank that passes columns other than its own is refused before any collective
    o = dataclasses.replace(opts, num_partitions=16, hash_rate=16)
    cols = zkl.sharded.DeviceColumns(ctx, trace, 0, 32)
    try:
        with pytest.raises(zkl.capi.ZklError, match="zkl_shard_columns"):
            zkl.sharded.prove_segment_sharded(ctx, lib.zkl_comm_loopback(), cols.ptr, 0, 32, trace.shape[0], trace.shape[1], o, air_pi)
    finally:
        cols.close()
