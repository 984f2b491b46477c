#!/usr/bin/env python
"""What the box's NCCL gives the column-sharded prover: all_to_all_single / all_gather_into_tensor / all_reduce bandwidth on large
device buffers, one process per GPU (torchrun). Run with NCCL_DEBUG=INFO to see the transport (P2P/NVLink, NVLS, SHM)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nbytes = mb << 20
send = torch.ones(nbytes, dtype=torch.uint8, device=dev)
recv = torch.empty(nbytes, dtype=torch.uint8, device=dev)
big = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.time()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.time() - t0) / reps


res = {}
res["all_to_all_single"] = timed(lambda: dist.all_to_all_single(recv, send))
res["all_gather_into_tensor"] = timed(lambda: dist.all_gather_into_tensor(big, send))
res["all_reduce_int64"] = timed(lambda: dist.all_reduce(send.view(torch.int64)))
if rank == 0:
    a2a = nbytes * (world - 1) / world / res["all_to_all_single"] / 1e9
    ag = nbytes * (world - 1) / res["all_gather_into_tensor"] / 1e9
    ar = 2 * nbytes * (world - 1) / world / res["all_reduce_int64"] / 1e9
    print(f"world={world} buffer={mb} MiB: all_to_all {res['all_to_all_single'] * 1e3:.2f} ms ({a2a:.1f} GB/s sent per rank), "
          f"all_gather {res['all_gather_into_tensor'] * 1e3:.2f} ms ({ag:.1f} GB/s received per rank), "
          f"all_reduce {res['all_reduce_int64'] * 1e3:.2f} ms ({ar:.1f} GB/s bus)", flush=True)
    can = [torch.cuda.can_device_access_peer(0, j) for j in range(1, torch.cuda.device_count())]
    print("peer access from device 0:", can, flush=True)
dist.destroy_process_group()
