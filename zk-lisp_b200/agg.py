"""The aggregation layer on top of the step proofs: `ZlAggAir` and its trace (SURVEY.md §8 "agg prover hooks", §8f N2).

Reference: zk-lisp-proof-winterfell/src/agg/layout.rs:97-180 (the 31 columns), agg/air.rs:22-332 (24 transition constraints, 5
assertions, one periodic column of cycle n), agg/trace.rs:155-247 (`build_agg_trace_from_transcripts`), :261-700
(`build_agg_trace_core`: one row per child, padded to a power of two >= 8, chain errors, accumulators), :1436-1685
(`sample_fri_fold_child`), prove.rs:629-702 (`prove_agg_proof`: options, field extension).

The trace is 31 x 16 elements for a rollup-bench batch: it is built on the host from the children (Python integers); proving it is
the same device pipeline as a step proof with `air_kind = ZKL_AIR_AGG` (and `field_extension = 2` when 128 bits are asked for).
The reference's error-aggregate columns (DEEP vs FRI layer 0, layer-1 consistency, folding along the query paths, Merkle roots) are
weighted sums of per-query errors; here every child is VERIFIED first (replay.py: openings, DEEP, FRI), which makes each of those
sums zero whatever the weights, and a child that fails raises instead of yielding a trace that violates the AIR.
"""
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import capi
from . import replay as R
from .air import ProofOptions, context_to_elements, select_partitions_for_trace
from .agg_child import ChildTranscript, child_transcript_from_step
from .field import GENERATOR, P, fold_bytes32_to_fe, inv, root_of_unity
from .recursion import AggAirPublicInputs
from .step import InvalidInput, StepProof, children_root

COLUMNS = ["ok", "v0_sum", "v1_sum", "vnext_sum", "fri_v0_child", "fri_v1_child", "fri_vnext_child", "fri_alpha_child", "fri_x0_child",
           "fri_x1_child", "fri_q1_child", "comp_sum", "alpha_div_zm_sum", "map_l0_sum", "final_llast_sum", "r", "alpha", "beta", "gamma",
           "seg_first", "trace_root_err", "constraint_root_err", "v_units_acc", "v_units_child", "child_count_acc", "vm_chain_err",
           "ram_u_chain_err", "ram_s_chain_err", "rom_chain_err_0", "rom_chain_err_1", "rom_chain_err_2"]   # layout.rs:97-180
COL = {name: i for i, name in enumerate(COLUMNS)}
WIDTH = len(COLUMNS)
MIN_AGG_TRACE_ROWS = 8            # agg/trace.rs:35
NUM_TRANSITION = 24               # agg/air.rs:69-94
assert WIDTH == capi.AGG_WIDTH and NUM_TRANSITION == capi.AGG_TRANSITIONS


# ---- ZlAggAir on Python integers -------------------------------------------------------------------------------------------
def transition_degrees(n: int) -> List[Tuple[int, int]]:
    """agg/air.rs:69-94 as (base degree, number of cycle-n periodic factors)."""
    return [(1, 0), (2, 1)] + [(1, 0)] * 9 + [(1, 1)] + [(1, 0)] * 12


def evaluate_transition(cur: Sequence[int], nxt: Sequence[int], is_last: int) -> List[int]:
    """agg/air.rs:112-272."""
    c = COL
    nl = (1 - is_last) % P
    out = [cur[c["ok"]],
           nl * (nxt[c["v_units_acc"]] - (cur[c["v_units_acc"]] + cur[c["v_units_child"]] * cur[c["seg_first"]])),
           cur[c["trace_root_err"]], cur[c["constraint_root_err"]]]
    out += [nl * (nxt[c[k]] - cur[c[k]]) for k in ("r", "alpha", "beta", "gamma", "v0_sum", "v1_sum", "vnext_sum")]
    out.append(nl * (nxt[c["child_count_acc"]] - (cur[c["child_count_acc"]] + cur[c["seg_first"]])))
    v0, v1, vn, al, x0, x1 = (cur[c[k]] for k in ("fri_v0_child", "fri_v1_child", "fri_vnext_child", "fri_alpha_child", "fri_x0_child", "fri_x1_child"))
    out.append(vn * (x1 - x0) - (v1 * (al - x0) - v0 * (al - x1)))
    out.append(vn - cur[c["fri_q1_child"]])
    out += [cur[c[k]] for k in ("comp_sum", "alpha_div_zm_sum", "map_l0_sum", "final_llast_sum", "vm_chain_err", "ram_u_chain_err",
                                "ram_s_chain_err", "rom_chain_err_0", "rom_chain_err_1", "rom_chain_err_2")]
    return [v % P for v in out]


def assertions(n: int, v_units_total: int, children_count: int) -> List[Tuple[int, int, int]]:
    """agg/air.rs:274-305."""
    last = n - 1
    return [(COL["ok"], 0, 0), (COL["v_units_acc"], 0, 0), (COL["v_units_acc"], last, v_units_total % P),
            (COL["child_count_acc"], 0, 0), (COL["child_count_acc"], last, children_count % P)]


def _np2(x):
    r = 1
    while r < x:
        r <<= 1
    return r


@dataclass
class AggAir:
    """What `ZlAggAir::new(info, pub_inputs, options)` derives (agg/air.rs:32-113) plus AirContext's two derived numbers [WF-mem]."""
    n: int
    agg_pi: AggAirPublicInputs
    opts: ProofOptions

    def __post_init__(self):
        deg = transition_degrees(self.n)
        self.num_transition = len(deg)
        self.ce_blowup = max(max(_np2(b + c) for b, c in deg), 2)
        highest = max(b * (self.n - 1) + c * (self.n - 1) for b, c in deg)       # a cycle-n periodic column has degree n - 1
        self.num_composition_cols = max((highest - (self.n - 1) + self.n - 1) // self.n, 1)
        self.assertions = assertions(self.n, self.agg_pi.v_units_total, self.agg_pi.children_count)
        self.num_constraints = self.num_transition + len(self.assertions)

    def seed_elements(self) -> List[int]:
        """Context::to_elements() ++ AggAirPublicInputs::to_elements() (agg/pi.rs:176-217)."""
        return context_to_elements(WIDTH, self.n, self.num_constraints, self.opts) + self.agg_pi.to_elements()

    def to_c(self) -> capi.AirParams:
        """Through the library's own builder (`zkl_agg_air_params`)."""
        a = capi.AirParams()
        self._assertion_array = (capi.Assertion * 5)()
        if capi.load_library().zkl_agg_air_params(self.n, self.agg_pi.v_units_total, self.agg_pi.children_count, a, self._assertion_array):
            raise ValueError("zkl_agg_air_params: the aggregation trace length must be a power of two >= 8")
        assert (a.num_transition, a.ce_blowup, a.num_composition_cols, a.num_assertions) == \
               (self.num_transition, self.ce_blowup, self.num_composition_cols, len(self.assertions))
        return a


def agg_proof_options(queries: int, blowup: int, grind: int, min_security_bits: int, n: int) -> ProofOptions:
    """prove.rs:629-681: at least 16 queries, quadratic extension from 128 bits up, partitions by trace shape."""
    num_partitions, hash_rate = select_partitions_for_trace(WIDTH, n)
    return ProofOptions(num_queries=max(queries, 16), blowup=blowup, grinding=grind, num_partitions=num_partitions, hash_rate=hash_rate,
                        field_extension=2 if min_security_bits >= 128 else 1)


# ---- the aggregation trace ---------------------------------------------------------------------------------------------------
def sample_fri_fold_child(tx: ChildTranscript, folding: int = 2):
    """agg/trace.rs:1436-1685: the first opened FRI coset of layer 0 (v0, v1, alpha, x0, x1), its fold, and the value layer 1 holds
    for the same path. Returns (v0, v1, vnext, alpha, x0, x1, q1)."""
    zero = (0,) * 7
    if not tx.fri_layers or tx.num_queries == 0:
        return zero
    if folding != 2:
        raise InvalidInput("AggTrace currently supports only FRI folding factor 2 in sample_fri_fold_child")
    if len(tx.fri_layers) < 2:
        raise InvalidInput("AggTrace requires at least two FRI layers in sample_fri_fold_child when num_queries > 0")
    fs = tx.compact.fs_challenges
    if fs is None:
        return zero
    if len(fs.query_positions) != tx.num_queries:
        raise InvalidInput("FS query_positions length is inconsistent with child transcript num_queries")
    N = tx.compact.meta.m * tx.compact.meta.rho
    if N == 0 or N & (N - 1):
        raise InvalidInput("AggTrace LDE domain size must be a non-zero power of two in sample_fri_fold_child")
    base_g = root_of_unity(N.bit_length() - 1)
    folded = R.fold_positions(fs.query_positions, N, 2)
    layer0, layer1 = tx.fri_layers[0], tx.fri_layers[1]
    if not folded or not layer0:
        return zero
    if len(layer0) != len(folded):
        raise InvalidInput("FRI folded positions count is inconsistent with transcript layer 0 values in sample_fri_fold_child")
    x0 = pow(base_g, folded[0], P) * GENERATOR % P
    x1 = x0 * pow(base_g, N // 2, P) % P
    v0, v1 = layer0[0]
    if not fs.fri_alphas:
        return zero
    alpha = fs.fri_alphas[0]
    den = (x1 - x0) % P
    if den == 0:
        raise InvalidInput("FRI coset has zero denominator in sample_fri_fold_child")
    vnext = (v1 * (alpha - x0) - v0 * (alpha - x1)) % P * inv(den) % P
    n1 = N // 2
    folded1 = R.fold_positions(folded, n1, 2)
    if len(layer1) != len(folded1):
        raise InvalidInput("FRI folded positions count is inconsistent with transcript layer 1 values in sample_fri_fold_child")
    row_len = n1 // 2
    coset, elem = folded[0] % row_len, folded[0] // row_len
    if coset not in folded1:
        raise InvalidInput("FRI sample coset index not found in folded positions at depth 1 in sample_fri_fold_child")
    q1 = layer1[folded1.index(coset)][elem]
    return v0 % P, v1 % P, vnext, alpha % P, x0, x1, q1 % P


def build_agg_trace(hasher, agg_pi: AggAirPublicInputs, steps: Sequence[StepProof], transcripts: Sequence[ChildTranscript] = None) -> np.ndarray:
    """`build_agg_trace_from_transcripts` (agg/trace.rs:155-247 over :261-700): (31, n, 2) uint64, column-major. `hasher` re-hashes
    the children's openings (a GPU Context, or any object with the batched hasher entry points)."""
    if not steps:
        raise InvalidInput("AggTrace requires at least one child transcript")
    k = len(steps)
    transcripts = transcripts or [child_transcript_from_step(hasher, s) for s in steps]
    kids = [t.compact for t in transcripts]
    if agg_pi.children_count != k:
        raise InvalidInput("AggAirPublicInputs.children_count must match number of children")
    if len(agg_pi.children_ms) != k:
        raise InvalidInput("AggAirPublicInputs.children_ms length must match number of children")
    if any(c.suite_id != agg_pi.suite_id for c in kids):
        raise InvalidInput("AggAirPublicInputs.suite_id must match suite_id of all children")
    totals = {c.segments_total for c in kids}
    if 0 in totals:
        raise InvalidInput("AggTrace requires child.segments_total > 0")
    if any(c.segment_index >= c.segments_total for c in kids):
        raise InvalidInput("AggTrace segment_index must satisfy 0 <= idx < segments_total")
    if len(totals) != 1:
        raise InvalidInput("AggTrace requires a uniform segments_total across children")
    t = totals.pop()
    if t > 1:
        if t != k:
            raise InvalidInput("AggTrace requires a complete contiguous segment chain in batch")
        if sorted(c.segment_index for c in kids) != list(range(k)):
            raise InvalidInput("AggTrace segment indices must form [0..children_count) without gaps")
    pm, pq = agg_pi.profile_meta, agg_pi.profile_queries
    for c in kids:
        if (c.meta.rho, c.meta.o, c.meta.lambda_bits, c.meta.pi_len) != (pm.rho, pm.o, pm.lambda_bits, pm.pi_len):
            raise InvalidInput("AggAirPublicInputs.profile_meta is inconsistent with child StepMeta")
        if c.meta.q != pq.num_queries:
            raise InvalidInput("AggAirPublicInputs.profile_queries.num_queries is inconsistent with child meta.q")
    if children_root(agg_pi.suite_id, [(c.step_digest, c.trace_root) for c in kids]) != agg_pi.children_root:
        raise InvalidInput("AggAirPublicInputs.children_root is inconsistent with child commitments")
    for m, c in zip(agg_pi.children_ms, kids):
        if m == 0:
            raise InvalidInput("AggAirPublicInputs.children_ms entries must be non-zero")
        if m != c.meta.m:
            raise InvalidInput("AggAirPublicInputs.children_ms entry does not match child meta.m")
    if agg_pi.v_units_total != sum(c.meta.v_units for c in kids):
        raise InvalidInput("AggAirPublicInputs.v_units_total must equal sum of child meta.v_units")
    # every opening, the DEEP composition and the FRI folding of every child: the reference folds these checks into error columns
    # (trace_root_err, constraint_root_err, comp_sum, alpha_div_zm_sum, map_l0_sum, final_llast_sum) that must be zero
    fss = [c.fs_challenges for c in kids]
    if all(f is not None for f in fss):
        R.verify_children_openings(hasher, list(steps), fss)
        for s, f in zip(steps, fss):
            R.verify_child_deep_and_fri(s, f)
    n = _np2(max(k, MIN_AGG_TRACE_ROWS))
    tr = [[0] * n for _ in range(WIDTH)]
    f32 = fold_bytes32_to_fe
    v_acc = 0
    prev = None
    for i, (c, tx) in enumerate(zip(kids, transcripts)):
        first, last = i == 0, i + 1 == k
        vm_in, vm_out = f32(c.state_in_hash), f32(c.state_out_hash)
        ru_in, ru_out, rs_in, rs_out = f32(c.ram_gp_unsorted_in), f32(c.ram_gp_unsorted_out), f32(c.ram_gp_sorted_in), f32(c.ram_gp_sorted_out)
        rom_in0, rom_out0 = f32(c.rom_s_in[0]), f32(c.rom_s_out[0])
        if first or prev is None:
            errs = [vm_in - f32(agg_pi.vm_state_initial), ru_in - f32(agg_pi.ram_gp_unsorted_initial), rs_in - f32(agg_pi.ram_gp_sorted_initial),
                    rom_in0 - f32(agg_pi.rom_s_initial[0])]
        else:
            errs = [vm_in - prev[0], ru_in - prev[1], rs_in - prev[2], rom_in0 - prev[3]]
        if last:
            errs[0] += vm_out - f32(agg_pi.vm_state_final)
            errs[1] += ru_out - f32(agg_pi.ram_gp_unsorted_final)
            errs[2] += rs_out - f32(agg_pi.ram_gp_sorted_final)
            errs[3] += rom_out0 - f32(agg_pi.rom_s_final[0])      # lanes 1, 2 of the ROM state are per level, not chained (trace.rs:523-543)
        tr[COL["seg_first"]][i] = 1
        tr[COL["v_units_child"]][i] = c.meta.v_units % P
        tr[COL["v_units_acc"]][i] = v_acc
        tr[COL["child_count_acc"]][i] = i
        for name, e in zip(("vm_chain_err", "ram_u_chain_err", "ram_s_chain_err", "rom_chain_err_0"), errs):
            tr[COL[name]][i] = e % P
        sample = sample_fri_fold_child(tx, agg_pi.profile_fri.folding_factor)
        for name, v in zip(("fri_v0_child", "fri_v1_child", "fri_vnext_child", "fri_alpha_child", "fri_x0_child", "fri_x1_child", "fri_q1_child"), sample):
            tr[COL[name]][i] = v
        v_acc = (v_acc + c.meta.v_units) % P
        prev = (vm_out, ru_out, rs_out, rom_out0)
    for r in range(k, n):                       # padding rows carry the accumulators
        tr[COL["v_units_acc"]][r] = v_acc
        tr[COL["child_count_acc"]][r] = k
    out = np.zeros((WIDTH, n, 2), dtype=np.uint64)
    for c in range(WIDTH):
        for r in range(n):
            out[c, r, 0], out[c, r, 1] = tr[c][r] & 0xFFFFFFFFFFFFFFFF, tr[c][r] >> 64
    return out


def check_trace(trace: np.ndarray, air: AggAir) -> None:
    """Every transition constraint on the trace domain and every assertion (what Winterfell's debug-mode validation would catch)."""
    n = trace.shape[1]
    rows = [[int(trace[c, r, 0]) | (int(trace[c, r, 1]) << 64) for c in range(WIDTH)] for r in range(n)]
    for r in range(n - 1):      # the last step is exempt (one transition exemption)
        res = evaluate_transition(rows[r], rows[r + 1], 0)
        bad = [i for i, v in enumerate(res) if v]
        if bad:
            raise InvalidInput(f"aggregation trace violates constraint {bad[0]} at row {r}")
    for col, row, val in air.assertions:
        if rows[row][col] != val:
            raise InvalidInput(f"aggregation trace violates the assertion on column {col} at row {row}")


def prove_agg_proof(ctx, agg_pi: AggAirPublicInputs, steps: Sequence[StepProof], queries: int = 64, blowup: int = 16, grind: int = 16,
                    min_security_bits: int = 128) -> bytes:
    """prove.rs:629-702 on the device: build the aggregation trace from the children, prove it under `ZlAggAir`."""
    import ctypes as C
    from .prover import ints_to_fe
    from .verify import conjectured_security_bits
    trace = build_agg_trace(ctx, agg_pi, steps)
    n = trace.shape[1]
    opts = agg_proof_options(queries, blowup, grind, min_security_bits, n)
    if min_security_bits >= 64:                  # prove.rs:664-681 (zk-lisp's own estimate: no LDE-size term)
        q_sec = (blowup.bit_length() - 1) * opts.num_queries
        bits = min(min(128 * opts.field_extension, q_sec + (grind if q_sec >= 80 else 0)) - 1, 128)
        if bits < min_security_bits:
            raise InvalidInput("aggregation prover options do not achieve requested min_security_bits; increase --queries/--blowup/--grind or lower --security-bits")
    del conjectured_security_bits
    air = AggAir(n, agg_pi, opts)
    check_trace(trace, air)
    ap, o = air.to_c(), opts.to_c()
    seed = ints_to_fe(air.seed_elements())
    buf = np.empty(1 << 20, dtype=np.uint8)
    cap = C.c_size_t(buf.shape[0])
    summ = capi.ProofSummary()
    ctx._check(ctx.lib.zkl_prove_segment(ctx.h, trace.ctypes.data, WIDTH, n, C.byref(ap), seed.ctypes.data, seed.shape[0], C.byref(o),
                                         buf.ctypes.data, C.byref(cap), C.byref(summ)))
    return buf[:cap.value].tobytes()
