"""Column-sharded trace commitment (zk-lisp_b200/sharded.py) on the GPU: the NCCL flow with a world of ONE rank must give
the root of the single-GPU hook (`zkl_trace_lde_commit`) and of the CPU oracle. The two-rank exchange itself is
exercised by tools/sharded_commit_check.py under torchrun (2 GPUs) and, for the index logic, by the gloo tests."""
import os
import socket

import numpy as np
import pytest

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
