#!/usr/bin/env python
"""bench.py — rollup-bench proving throughput on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: the 16 execution segments of rollup-bench
(n = 4096 trace rows, W = 204 columns, blowup 16, 64 queries, grinding 16; README.md:186-194 of the
reference) are proven, i.e. 65 536 trace rows per step per GPU. The 16 segments belong to ONE
program (one program id / commitment, consecutive pc ranges), like the reference run. Segments are independent
(prove.rs:1020-1048), so with N GPUs every rank proves its own 16 segments (weak scaling, no collective on
the data path); the strong-scaling figure (ONE 16-segment job split over the N GPUs = BASELINE's "prove wall-time
at 1/2/4/8") is measured in the same run and reported under `prove_wall_ms`. The synthetic traces satisfy the
zk-lisp VM AIR (tests/synth.py), so every proof is a real, verifiable proof.

  value    : trace rows / s, traces already resident in HBM when the timed region starts
  e2e      : same through the C-ABI call with HOST buffers (pinned), H2D of the trace and D2H of the proof
             bytes inside the timed region. The per-segment AIR instance (assertions, degrees, coin seed) is built
             inside every call from the public inputs (zkl_prove_segment_pi); nothing about a segment is cached.
  e2e_cold : same, and additionally every prover object is new and the per-program constants (blake3-derived
             suites, Cauchy MDS) are re-derived for every segment: a first-ever prove through the public API.
Timing: CUDA events on the proving streams (max over streams, then max over ranks), barrier +
synchronize on both sides; wall clock is reported beside it as a cross-check.

--impl reference: the CPU restatement of the same pipeline (oracle/, the reference itself is Rust and cannot be built
here) at the SAME configuration — n = 4096, grind 16, one thread per segment, min(16, cores) in flight.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SEGMENTS, ROWS, QUERIES, BLOWUP, GRIND = 16, 4096, 64, 16, 16
PROGRAM_SEED = 0x5A4B4C
METRIC, UNIT = "rollup-bench prove throughput (trace rows/s)", "trace rows/s"
WORKLOAD = "rollup-bench: 16 segments x (n=4096, W=204, blowup=16, q=64, grind=16) of one program, synthetic AIR-valid traces"
# algorithmic work of the dominant kernel (trace row hashing, k_tc_hash_rows), SURVEY.md §8(d):
# W=204 -> 1 + 102 absorbs -> 11 permutations per LDE row; one permutation = 27 rounds; per round the 12x12 MDS
# product over f128 runs on the tensor cores as a u8 x u8 -> s32 GEMM tile: 192 output byte-rows x 384 cube bytes
# per sponge (zk-lisp_b200/csrc/tcperm.cuh) = 73 728 MACs, plus 12 cubes (24 128x128-bit products) on the integer pipe.
PERMS_PER_ROW = 11
TENSOR_OPS_PER_PERM = 27 * 2 * 192 * 384
FIELD_PRODUCTS_PER_PERM = 27 * (24 + 144)
WIDE_MACS_PER_FIELD_PRODUCT = 16      # 4x4 32-bit limb products of one 128x128-bit multiplication (reduction not counted)
# dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the ncu --set full captures committed under profiles/
# (not measured in this run: ncu replays kernels, a bench run under it is not a bench run)
TRAFFIC = {
    "k_tc_hash_rows": (218.7e6, "profiles/prof_tc_hash_rows_r02_details.txt: 214.1 MB read + 4.6 MB written; algorithmic 16*204*65536 = 213.9 MB read + 1 MB written"),
    "k_tc_lde4096_v2": (172.0e6, "profiles/prof_tc_lde4096_v2_r02_details.txt: 15.7 MB read + 156.3 MB written inside the kernel's window (the rest of the "
                                 "213.9 MB output is still in the 126 MB L2 when the kernel ends); algorithmic 227.3 MB"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="concurrent segment streams per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 16 segments per GPU per step; strong: the 16 segments of one rollup-bench split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bulk-threshold", type=int, default=0)
    ap.add_argument("--reference-sample", action="store_true", help="--impl reference on the bounded n=512 / grind 13 sample instead of the full configuration")
    ap.add_argument("--config", default="rollup-bench", choices=["rollup-bench", "x-synth"],
                    help="x-synth: ONE large segment (W = 219, partitions (16, 16)) column-sharded over the N GPUs (BASELINE.json configs[4]; strong scaling)")
    ap.add_argument("--no-large-segment", action="store_true", help="skip the 65 536-row single-segment leg (prove_wall_ms.one_65536_row_segment)")
    ap.add_argument("--x-logn", type=int, default=20, help="x-synth: log2 of the trace length (22 = BASELINE's size; the Python trace generator needs ~45 us per row)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def pin_rank_to_cores(local, world):
    """Each rank keeps to its own share of the host cores: the proving threads of 8 ranks otherwise migrate over all cores."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if world > 1 and per >= 1:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_segment(i, rows=ROWS, grind=GRIND):
    """Segment i of the synthetic rollup-bench program: its own data, the program's id, pc range [i * levels, (i + 1) * levels)."""
    import synth
    return synth.valid_segment(n=rows, queries=QUERIES, blowup=BLOWUP, grind=grind, seed=PROGRAM_SEED + i, start_level=(i % SEGMENTS) * (rows // 32),
                               program_seed=PROGRAM_SEED)


def cpu_reference_run(zkl, threads, rows, grind):
    """The CPU restatement of the same pipeline (oracle/), one thread per segment, `threads` segments in
    flight — the shape of the reference's rayon pool over segments (prove.rs:1020-1048)."""
    import oracle_lib
    oracle_lib.build()
    segs = [make_segment(i, rows, grind) for i in range(threads)]
    airs = [zkl.air.Air(t.shape[0], rows, pi, o) for t, pi, o in segs]
    inputs = [(a.to_c(), zkl.prover.ints_to_fe(a.seed_elements())) for a in airs]

    def work(i):
        trace, _, opts = segs[i]
        oracle_lib.prove_segment(trace, inputs[i][0], inputs[i][1], opts.to_c())

    t0 = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.time() - t0
    return threads * rows / dt, dt


def run_reference(args, zkl, rank, world):
    """bench.py --impl reference: the CPU port at the b200 arm's configuration. One step = min(16, cores) segments of
    n = 4096 / grind 16 in flight, one thread each (the reference's Winterfell is single-threaded per segment);
    a step is minutes of CPU time, so ONE step is timed whatever --steps says."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    full = not args.reference_sample
    rows, grind = (ROWS, GRIND) if full else (512, 13)
    threads = min(cores, SEGMENTS) if full else min(cores, 256)
    v, dt = cpu_reference_run(zkl, threads, rows, grind)
    sample = (f"{threads} of the step's {SEGMENTS} segments in flight (one thread each), n={rows}, W=204, blowup=16, q=64, grind={grind}: "
              + ("the b200 arm's configuration" if full else "bounded sample, grinding work per row kept proportional") + f"; {dt:.1f} s of wall time; "
              "CPU port of the reference pipeline (oracle/; the Rust reference cannot be built in this image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
        "ms_per_step": dt * 1e3 * (SEGMENTS / threads), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u128 (f128 field)", "data": "synthetic", "config": {"workload": WORKLOAD, "reference_sample": sample, "same_config": bool(full)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_x_synth(args, zkl, rank, world, local):
    """BASELINE.json configs[4]: ONE large synthetic zk-lisp VM segment (full 219-column layout, PartitionOptions (16, 16) = 14 column
    partitions), column-sharded over the N GPUs of the box (include/zkl_b200.h: zkl_prove_segment_sharded_*; NCCL through
    zk-lisp_b200/sharded.py: TorchComm). Strong scaling: the segment is the same at every N; N = 1 is the plain single-GPU prover.
    value: columns resident in HBM; e2e: every rank's columns start in pinned host memory and the proof bytes end on the host.
    Rank 0 generates the AIR-valid trace (tests/synth.py) and hands every rank its columns over NCCL before the timed region."""
    import dataclasses
    import torch
    import torch.distributed as dist
    import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the proving path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, W = 1 << args.x_logn, 219
    ctx = zkl.Context(local)
    lib = ctx.lib
    t0 = time.time()
    if rank == 0:
        trace, air_pi, opts = synth.valid_segment(n=n, queries=QUERIES, blowup=BLOWUP, grind=GRIND, seed=PROGRAM_SEED, full_layout=True, poseidon_levels=False)
        opts = dataclasses.replace(opts, num_partitions=16, hash_rate=16)
        assert trace.shape == (W, n, 2)
        meta = [air_pi, opts]
    else:
        trace, meta = None, [None, None]
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    air_pi, opts = meta
    gen_s = time.time() - t0
    if world > 1:
        first, cnt = zkl.sharded.shard_columns(lib, opts, W, rank, world)
    else:
        first, cnt = 0, W
    cols = torch.empty(cnt * n * 16, dtype=torch.uint8, device=dev)
    if rank == 0:
        for r in range(world - 1, -1, -1):
            f, c = zkl.sharded.shard_columns(lib, opts, W, r, world) if world > 1 else (0, W)
            t = torch.from_numpy(np.ascontiguousarray(trace[f:f + c]).view(np.uint8).reshape(-1))
            if r == 0:
                cols.copy_(t)
            else:
                stage = t.to(dev)
                dist.send(stage, dst=r)
                del stage
        del trace
    else:
        dist.recv(cols, src=0)
    torch.cuda.synchronize()
    host = torch.empty(cols.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(cols)
    comm = zkl.sharded.TorchComm(dist, dev) if world > 1 else None
    summ = zkl.capi.ProofSummary()
    prover = zkl.ZkProver(ctx, opts, air_pi)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def prove(from_host):
        if from_host:
            cols.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
        if world > 1:
            return zkl.sharded.prove_segment_sharded(ctx, comm.c, cols.data_ptr(), first, cnt, W, n, opts, air_pi, summary=summ), summ
        p = prover.prove(None, dev_ptr=C.c_void_p(cols.data_ptr()), shape=(W, n))
        return p, prover.last_summary

    def timed(from_host, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        launches0 = ctx.launches
        ev, wall, stages, digests = 0.0, 0.0, np.zeros(12), []
        for _ in range(steps):
            if world > 1:
                dist.barrier()
            t1 = time.time()
            proof, s = prove(from_host)
            wall += (time.time() - t1) * 1e3
            # device time of the proof: the library's own CUDA events on ITS stream, first stage mark to the last one (the collectives
            # run inside that window); torch.cuda.Event would only see torch's stream. The e2e leg is reported on the host wall clock.
            ev += float(sum(s.stage_ms[:9]))
            stages += np.array(list(s.stage_ms))
            digests.append(hashlib.blake2b(proof, digest_size=16).hexdigest())
        barrier()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ev, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), stages / steps, digests, len(proof), ctx.launches - launches0, clocks

    for _ in range(max(args.warmup, 1)):
        first_proof, _ = prove(False)
    ref_digest = hashlib.blake2b(first_proof, digest_size=16).hexdigest()
    if comm is not None:   # the warm-up proof also set up NCCL's peer connections: count the timed proofs only
        comm.bytes_moved, comm.calls = 0, 0
        comm.seconds = {k: 0.0 for k in comm.seconds}
    sampler = ClockSampler(local) if rank == 0 else None
    ev, wall, stages, digests, plen, launches, clocks = timed(False, args.steps, sampler)
    e_ev, e_wall, _, e_digests, _, _, _ = timed(True, args.steps)
    same = all(d == ref_digest for d in digests + e_digests)
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, (ref_digest, same, comm.bytes_moved, comm.calls, float(stages[10]), cnt, dict(comm.seconds)))
    else:
        every = [(ref_digest, same, 0, 0, float(stages[10]), cnt, {})]
    i8 = C.c_double()
    ctx._check(lib.zkl_tc_i8_peak(ctx.h, C.byref(i8)))
    if rank == 0:
        N = n * BLOWUP
        same_everywhere = all(e[0] == ref_digest and e[1] for e in every)
        # dominant kernel: partition hashing of the rank's columns, one permutation per (LDE row, 16-column partition)
        my_parts = -(-cnt // 16)
        hash_ms = float(stages[10])
        tensor_ops = N * my_parts * TENSOR_OPS_PER_PERM
        h2d = int(cols.numel())
        total_steps = 2 * args.steps   # both timed legs (the counters restart after the warm-up)
        line = {
            "metric": "x-synth single-segment prove throughput (trace rows/s)", "value": n * args.steps / (ev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u128 (f128 field, 4x u32 limbs)", "data": "synthetic",
            "config": {"workload": f"x-synth: ONE segment, n=2^{args.x_logn}, W=219, blowup=16, q=64, grind=16, partitions (16,16) -> 14 column partitions, "
                                   f"column-sharded over {world} GPU(s); AIR-valid synthetic VM trace", "columns_per_rank": [e[5] for e in every],
                       "l2": f"the rank's LDE alone is {cnt * N * 16 / 1e9:.1f} GB >> 126 MB L2 (no flush needed)",
                       "trace_generation_s": round(gen_s, 1)},
            "wall_ms_per_step": wall / args.steps,
            "e2e": {"value": n * args.steps / (e_wall * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": plen,
                    "wall_ms_per_step": e_wall / args.steps, "device_ms_per_step_without_h2d": e_ev / args.steps,
                    "what": "every rank's columns start in pinned host memory, H2D copy + column-sharded proof + proof bytes on the host; host wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "proof_bytes": plen, "proofs_identical_across_ranks_steps_and_legs": bool(same_everywhere),
            "proof_blake2b": ref_digest,
            "collectives": {"calls_per_proof": every[0][3] / total_steps, "bytes_sent_per_rank_per_proof": [e[2] / total_steps for e in every],
                            "host_ms_per_proof_rank0": {k: 1e3 * v / total_steps for k, v in every[0][6].items()},
                            "raw_rows_would_be": (world - 1) * cnt * N * 16 if world > 1 else 0},
            "clocks": clocks,
            "roofline": {"kernel": "k_tc_hash_rows (partition digests of the rank's columns: one sponge permutation per LDE row and 16-column partition)",
                         "bound": "tensor", "achieved": tensor_ops / (hash_ms * 1e-3) / 1e12, "peak": i8.value / 1e12, "unit": "TOP/s (int8 dense)",
                         "frac": tensor_ops / (hash_ms * 1e-3) / 1e12 / (i8.value / 1e12),
                         "peak_source": "measured in this run: zkl_tc_i8_peak", "traffic": None, "ms_per_launch": hash_ms,
                         "ms_per_launch_per_rank": [e[4] for e in every], "permutations_per_s": N * my_parts / (hash_ms * 1e-3)},
            "stage_ms_rank0": {lib.zkl_stage_name(i).decode(): float(stages[i]) for i in range(12)},
            "cpu_baseline": None,
            "cpu_baseline_note": "no CPU run exists at this size: Winterfell's boundary-divisor table is O(n^2) (SURVEY.md §8a row C); the rollup-bench line carries the CPU arm",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class Worker:
    """One CUDA stream (zkl_ctx) proving a share of the step's segments."""

    def __init__(self, zkl, device, segs, bulk_threshold):
        self.zkl, self.ctx = zkl, zkl.Context(device)
        if bulk_threshold:
            self.ctx.lib.zkl_set_bulk_threshold(self.ctx.h, bulk_threshold)
        self.items = []
        for trace, air_pi, opts in segs:
            prover = zkl.ZkProver(self.ctx, opts, air_pi)
            dev = C.c_void_p()
            self.ctx._check(self.ctx.lib.zkl_dev_alloc(self.ctx.h, trace.nbytes, C.byref(dev)))
            self.ctx._check(self.ctx.lib.zkl_dev_upload(self.ctx.h, dev, trace.ctypes.data, trace.nbytes))
            self.items.append((prover, trace, dev))
        self.ms, self.summaries, self.proof_bytes, self.digests = 0.0, [], 0, []
        self.active = len(self.items)

    def run(self, mode: str):
        """mode: "resident" (trace in HBM), "host" (host buffers through the C ABI), "cold" (host buffers, new prover objects and
        re-derived program constants for every segment)."""
        ms = C.c_float()
        self.summaries, self.proof_bytes = [], 0
        proofs = []
        self.ctx._check(self.ctx.lib.zkl_timer_start(self.ctx.h))
        for prover, trace, dev in self.items[:self.active]:
            if mode == "cold":
                prover = self.zkl.ZkProver(self.ctx, prover.options, prover.air_pi, cache_program=False)
            proof = prover.prove(None, dev_ptr=dev, shape=trace.shape[:2]) if mode == "resident" else prover.prove(trace)
            self.proof_bytes += len(proof)
            self.summaries.append(prover.last_summary)
            proofs.append(proof)
        self.ctx._check(self.ctx.lib.zkl_timer_stop(self.ctx.h, C.byref(ms)))
        self.ms = ms.value
        # outside the timed region: the same segment must give the same bytes in every step and through every entry point
        # (8 streams share the GPU: a racy kernel shows up here as proofs that differ from run to run)
        self.digests = [hashlib.blake2b(p, digest_size=16).hexdigest() for p in proofs]


def run_step(workers, mode):
    live = [w for w in workers if w.active]
    th = [threading.Thread(target=w.run, args=(mode,)) for w in live]
    t0 = time.time()
    [t.start() for t in th]
    [t.join() for t in th]
    wall = (time.time() - t0) * 1e3
    return max(w.ms for w in live), wall


def main():
    args = parse_args()
    rank, world, local = dist_env()
    import __graft_entry__ as graft
    zkl = graft.import_zkl()
    if args.impl == "reference":
        run_reference(args, zkl, rank, world)
        return
    if args.config == "x-synth":
        run_x_synth(args, zkl, rank, world, local)
        return
    host_cores = pin_rank_to_cores(local, world)
    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the proving path has no CPU fallback")
    device = local

    if args.scaling == "weak":
        my_ids = zkl.sharding.weak_scaling_segments(SEGMENTS, rank)
    else:
        my_ids = zkl.sharding.segments_for_rank(SEGMENTS, rank, world)
    segs = [make_segment(i) for i in my_ids]
    S = max(1, min(args.streams, len(segs)))
    workers = [Worker(zkl, device, segs[i::S], args.bulk_threshold) for i in range(S)]
    # pinned host buffers for the e2e leg
    for w in workers:
        for _, trace, _ in w.items:
            torch.cuda.cudart().cudaHostRegister(trace.ctypes.data, trace.nbytes, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mode, steps, sampler=None, check=True):
        barrier()
        if sampler:
            sampler.start()
        ev_ms, wall_ms = [], []
        launches0 = sum(w.ctx.launches for w in workers)
        nonlocal deterministic
        for _ in range(steps):
            e, wl = run_step(workers, mode)
            ev_ms.append(e); wall_ms.append(wl)
            if check and reference_digests and reference_digests[0]:
                deterministic = deterministic and [list(w.digests) for w in workers] == reference_digests
        barrier()
        clocks = sampler.stop() if sampler else None
        tot_ev, tot_wall = float(np.sum(ev_ms)), float(np.sum(wall_ms))
        if world > 1:
            t = torch.tensor([tot_ev, tot_wall], device="cuda", dtype=torch.float64)
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            per_rank[mode] = [round(float(x[0]) / steps, 2) for x in every]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot_ev, tot_wall = float(t[0]), float(t[1])
        return tot_ev, tot_wall, sum(w.ctx.launches for w in workers) - launches0, clocks

    per_rank = {}
    for _ in range(args.warmup):
        run_step(workers, "resident")
    reference_digests = [list(w.digests) for w in workers]   # from the last warm-up step
    deterministic = True
    # one sampler per job: every nvidia-smi query takes the driver's global lock, eight pollers on an eight-GPU box would perturb
    # the proving threads' launches on all GPUs
    sampler = ClockSampler(device) if rank == 0 else None
    ev_total, wall_total, launches, clocks = timed("resident", args.steps, sampler)
    stage = np.zeros(12)
    cnt = 0
    for w in workers:
        for s in w.summaries:
            stage += np.array(list(s.stage_ms)); cnt += 1
    stage /= max(cnt, 1)
    run_step(workers, "host")  # warm the host-buffer path
    e2e_ev, e2e_wall, _, _ = timed("host", args.steps)
    cold_steps = max(1, min(args.steps, 3))
    cold_ev, cold_wall, _, _ = timed("cold", cold_steps)
    h2d = sum(t.nbytes for w in workers for _, t, _ in w.items)
    d2h = sum(w.proof_bytes for w in workers)

    rows_local = len(segs) * ROWS * args.steps
    rows_total, _ = zkl.sharding.aggregate(rows_local, 0.0, dist if world > 1 else None, "cuda")
    value = rows_total / (ev_total * 1e-3)
    e2e_value = rows_total / (e2e_ev * 1e-3)
    cold_value = rows_total / args.steps * cold_steps / (cold_ev * 1e-3)

    # strong scaling: ONE 16-segment rollup-bench job split over the ranks (rank r proves segments r, r + world, ...), host buffers.
    # Fewer than 4 busy streams per GPU: the lowest-latency small-sponge kernels (a pool takes the 16-thread groups).
    strong_ms = None
    if args.scaling == "weak":
        share = len(zkl.sharding.segments_for_rank(SEGMENTS, rank, world))
        per = [share // S + (1 if i < share % S else 0) for i in range(S)]
        busy = sum(1 for p in per if p)
        for w, p in zip(workers, per):
            w.active = p
            w.ctx._check(w.ctx.lib.zkl_set_coop_path(w.ctx.h, 2 if busy < 4 else 0))
        run_step(workers, "host")
        s_ev, s_wall, _, _ = timed("host", max(3, min(args.steps, 10)), check=False)
        strong_ms = {"events": s_ev / max(3, min(args.steps, 10)), "wall": s_wall / max(3, min(args.steps, 10)), "segments_per_gpu": share}
        for w in workers:
            w.active = len(w.items)
            w.ctx._check(w.ctx.lib.zkl_set_coop_path(w.ctx.h, 0))

    # stream-exclusive pass (nothing else on the GPU) for the per-kernel roofline numbers
    barrier()
    for w in workers:
        w.active = 0
    workers[0].active = 1
    workers[0].ctx._check(workers[0].ctx.lib.zkl_set_coop_path(workers[0].ctx.h, 2))  # a lone prover: lowest-latency small-sponge kernels
    workers[0].run("resident")
    excl = np.array(list(workers[0].summaries[0].stage_ms))
    for _ in range(2):                                  # median of three lone proofs
        workers[0].run("resident")
        excl = np.vstack([excl.reshape(-1, 12), np.array(list(workers[0].summaries[0].stage_ms))])
    excl = np.median(excl, axis=0)
    lone_wall = []
    for _ in range(3):
        t0 = time.time(); workers[0].run("host"); lone_wall.append((time.time() - t0) * 1e3)
    # measured peaks of the two pipes the sponge kernels run on (this GPU, this run)
    i8 = C.c_double(); imad = C.c_double()
    workers[0].ctx._check(workers[0].ctx.lib.zkl_tc_i8_peak(workers[0].ctx.h, C.byref(i8)))
    workers[0].ctx._check(workers[0].ctx.lib.zkl_imad_peak(workers[0].ctx.h, C.byref(imad)))
    barrier()
    # one LARGE segment on the same stream (with the O(n^2) boundary table gone, fewer and larger segments are an
    # operating point the reference cannot reach): 65 536 rows = 16x the reference's segment cap, every transform on the tensor cores
    large = None
    if rank == 0 and not args.no_large_segment:
        import synth
        ltrace, lpi, lopts = synth.valid_segment(n=65536, queries=QUERIES, blowup=BLOWUP, grind=GRIND, seed=PROGRAM_SEED + 99, program_seed=PROGRAM_SEED)
        lprover = zkl.ZkProver(workers[0].ctx, lopts, lpi)
        lproof = lprover.prove(ltrace)                                   # warm-up: tables of this size
        lms = []
        for _ in range(2):
            lproof = lprover.prove(ltrace)
            lms.append(float(sum(lprover.last_summary.stage_ms[:9])))
        large = {"rows": 65536, "width": int(ltrace.shape[0]), "device_ms": min(lms), "rows_per_s": 65536 / (min(lms) * 1e-3), "proof_bytes": len(lproof),
                 "constraints_violated": int(lprover.last_summary.constraints_violated), "what": "host trace -> proof bytes on ONE stream; device time of the proof"}
        del ltrace, lproof

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
        N = ROWS * BLOWUP
        hash_ms, lde_ms = float(excl[10]), float(excl[9])
        perms = N * PERMS_PER_ROW
        lde_bytes = 16 * 204 * ROWS * (1 + BLOWUP)
        tensor_ops = perms * TENSOR_OPS_PER_PERM
        int8_peak = i8.value / 1e12
        products_per_s = perms * FIELD_PRODUCTS_PER_PERM / (hash_ms * 1e-3)
        roofline = {
            "kernel": "k_tc_hash_rows (trace row hashing: 11 sponge permutations per LDE row, MDS product on tcgen05 kind::i8)",
            "bound": "tensor",
            "achieved": tensor_ops / (hash_ms * 1e-3) / 1e12, "peak": int8_peak, "unit": "TOP/s (int8 dense)",
            "frac": tensor_ops / (hash_ms * 1e-3) / 1e12 / int8_peak,
            "peak_source": "measured in this run: zkl_tc_i8_peak, a saturating tcgen05.mma kind::i8 loop (M128 N256 K32 from shared memory, one CTA per SM)",
            "peak_cublas_bf16_x2": 2.0 * peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
            "traffic": TRAFFIC["k_tc_hash_rows"][0], "traffic_source": TRAFFIC["k_tc_hash_rows"][1],
            "ms_per_launch": hash_ms, "ms_per_launch_in_timed_region": float(stage[10]),
            "permutations_per_s": perms / (hash_ms * 1e-3),
            "field_products_per_s": products_per_s,
            "integer_pipe": {
                "imad_wide_peak_macs_per_s": imad.value,
                "peak_source": "measured in this run: zkl_imad_peak, dependency-free IMAD.WIDE.U32 chains on every SM",
                "field_products_per_s_if_all_on_imad_pipe": imad.value / WIDE_MACS_PER_FIELD_PRODUCT,
                "achieved_over_imad_only_bound": products_per_s / (imad.value / WIDE_MACS_PER_FIELD_PRODUCT),
                "note": "144 of the 168 field products of a round run on the tensor cores; a ratio above 1 is that win: more field "
                        "products per second than the wide-multiply pipe alone could deliver (reductions not even counted)",
            },
            "co_bound": "the S-box cubes (IMAD.WIDE.U32, 5.2 clk per warp instruction) and the accumulator carry propagation stay on the CUDA cores",
            "timing": "ms_per_launch is taken on an otherwise idle GPU right after the timed region; inside it the kernel shares the SMs with the other streams",
        }
        roofline_lde = {
            "kernel": "k_tc_lde4096_v2 (trace coset LDE, 204 columns x 16 cosets of 4096 rows: three radix-16 stages as u8 GEMM tiles on tcgen05, "
                      "coset shift folded into per-coset byte matrices, columns fetched by TMA bulk copies; one pass over HBM)",
            "bound": "hbm", "achieved": lde_bytes / (lde_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": lde_bytes / (lde_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
            "traffic": TRAFFIC["k_tc_lde4096_v2"][0], "traffic_source": TRAFFIC["k_tc_lde4096_v2"][1], "ms_per_launch": lde_ms,
        }
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = min(os.cpu_count() or 1, 256)
            v, dt = cpu_reference_run(zkl, cores, 512, 13)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"bounded sample: {cores} segments in flight, n=512, W=204, blowup=16, q=64, grind=13 ({dt:.1f} s of wall time); "
                             "the full configuration (n=4096, grind 16) is what `--impl reference` times"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ev_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u128 (f128 field, 4x u32 limbs)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "segments_per_gpu_per_step": len(segs), "streams_per_gpu": S, "host_cores_per_rank": host_cores,
                       "l2": "per-step working set 16 x 250 MB LDE matrices >> 126 MB L2 (no flush needed)",
                       "published_reference": "109.567 s on Apple M3 Max for the same 65536 rows (BASELINE.md), different hardware"},
            "wall_ms_per_step": wall_total / args.steps,
            "ms_per_step_per_rank": per_rank.get("resident"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ev / args.steps, "wall_ms_per_step": e2e_wall / args.steps,
                    "per_segment_air_build": "inside every call (zkl_prove_segment_pi): assertions, degrees and coin seed from the public inputs; no per-segment cache",
                    "host_air_build_ms_per_segment": float(stage[11])},
            "e2e_cold": {"value": cold_value, "unit": UNIT, "ms_per_step": cold_ev / cold_steps, "wall_ms_per_step": cold_wall / cold_steps, "steps": cold_steps,
                         "what": "host buffers, a new prover object per segment, per-program constants re-derived for every segment (zkl_program_derive), "
                                 "AIR instance built inside the call"},
            "gpu_launches": int(launches),
            "proofs_identical_across_steps_and_entry_points": bool(deterministic),
            "prove_wall_ms": {"step_of_16_segments": ev_total / args.steps, "one_segment_alone_on_the_gpu": float(np.sum(excl[:9])),
                              "one_segment_alone_host_wall": float(np.min(lone_wall)),
                              "one_65536_row_segment": large,
                              "rollup_bench_over_N_gpus": strong_ms,
                              "rollup_bench_over_N_gpus_what": "ONE job of 16 segments split over the N GPUs (strong scaling), host buffers, max over ranks"},
            "clocks": clocks,
            "roofline": roofline, "roofline_lde": roofline_lde,
            "stage_ms_per_segment": {zkl.capi.load_library().zkl_stage_name(i).decode(): float(stage[i]) for i in range(12)},
            "stage_ms_one_segment_alone": {zkl.capi.load_library().zkl_stage_name(i).decode(): float(excl[i]) for i in range(12)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
