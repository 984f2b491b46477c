"""Host verifier of a step proof — the counterpart of `winterfell::verify` as the reference calls it (prove.rs:888-893, CLI `verify`),
restated on Python integers with the hashing on the device hasher: transcript replay (agg/fs.rs:37-245), every Merkle opening,
DEEP composition and FRI folding (agg/trace.rs:752-1193), and the out-of-domain consistency check — the constraints evaluated at z
from the OOD frame (air_eval.py), divided by their divisors and combined with the composition coefficients, must equal the
composition columns recombined at z (winter-verifier `evaluate_constraints` / `perform_verification` [WF-mem]).
One proof costs ~40 small device batches and a few thousand integer products; it is a checker for operators and tests, not a hot path.
"""
from typing import List

from . import air as A
from . import air_eval, params
from . import replay as R
from .field import P, be_from_le8, inv, root_of_unity
from .step import InvalidInput, StepProof


def periodic_at(n: int, x: int) -> List[int]:
    """The 31 cycle-32 periodic columns (vm/air/mod.rs:520-592, vm/schedule.rs:35-80; evaluated at x^(n/32)) and p_last (cycle n)."""
    y = pow(x, n // 32, P)
    w32 = root_of_unity(5)
    if pow(y, 32, P) == 1:                  # x on the trace domain: the Lagrange basis degenerates to the 0/1 schedule itself
        L = [int(y == pow(w32, p, P)) for p in range(32)]
    else:
        num = (pow(y, 32, P) - 1) * inv(32) % P
        L = [num * pow(w32, p, P) % P * inv((y - pow(w32, p, P)) % P) % P for p in range(32)]
    cols = [L[0]] + [L[1 + j] for j in range(27)] + [L[28], (L[29] + L[30] + L[31]) % P, L[31]]
    last = pow(root_of_unity(n.bit_length() - 1), n - 1, P)
    if pow(x, n, P) == 1:
        cols.append(int(x == last))
    else:
        cols.append(last * inv(n) % P * (pow(x, n, P) - 1) % P * inv((x - last) % P) % P)
    return cols


def air_of(step: StepProof, segment_feature_mask: int = 0) -> A.Air:
    """The AIR instance of a step proof, from its own bytes (agg/fs.rs:43-87). `segment_feature_mask` selects the segment's
    effective features when they differ from the program's (prove.rs:1082-1084); 0 = the program mask."""
    v = step.proof.inner
    from dataclasses import replace
    pi = replace(R.air_public_inputs_of(step), segment_feature_mask=segment_feature_mask or step.pi_core.feature_mask)
    return A.Air(v.main_width, v.trace_len, pi, R._options_of(v))


def ood_consistency(hasher, step: StepProof, air: A.Air):
    """(lhs, rhs, z): constraint side and composition side of the out-of-domain check; equal for a valid proof."""
    v = step.proof.inner
    n, W = v.trace_len, v.main_width
    seed = A.context_to_elements(W, n, v.num_constraints, R._options_of(v)) + R.air_public_inputs_of(step).to_elements()
    coin = R.Coin(hasher, seed)
    trace_roots, constraint_root, _ = v.parse_commitments()
    coin.reseed(trace_roots[0])
    cc = coin.draw_many(v.num_constraints)           # transition coefficients, then one per assertion in sorted order
    coin.reseed(constraint_root)
    z = coin.draw()
    fe = lambda b, i: int.from_bytes(b[16 * i:16 * i + 16], "little")
    C = len(v.ood_constraints) // 32
    if len(v.ood_trace) != 32 * W or C == 0 or W != air.width:
        raise InvalidInput("invalid OOD frame layout in FS replay")
    t_z, t_zg = [fe(v.ood_trace, c) for c in range(W)], [fe(v.ood_trace, W + c) for c in range(W)]
    h_z = [fe(v.ood_constraints, c) for c in range(C)]
    w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
    core = air.air_pi.core
    res = air_eval.evaluate_transition(air, t_z, t_zg, periodic_at(n, z), w0, w1, be_from_le8(core.merkle_root),
                                       rom_enabled=int(any(core.program_commitment)))
    T = len(res)
    if T + len(air.assertions) != v.num_constraints:
        raise InvalidInput("constraint count of the AIR does not match the proof context")
    g = root_of_unity(n.bit_length() - 1)
    zn = pow(z, n, P)
    lhs = sum(c * r for c, r in zip(cc, res)) % P * ((z - inv(g)) % P) % P * inv((zn - 1) % P) % P    # divisor (x^n - 1)/(x - g^(n-1))
    for k, (col, row, val) in enumerate(sorted(air.assertions, key=lambda t: (t[1], t[0]))):           # by step, then column
        lhs = (lhs + cc[T + k] * (t_z[col] - val) % P * inv((z - pow(g, row, P)) % P)) % P
    rhs = sum(pow(zn, j, P) * h for j, h in enumerate(h_z)) % P
    return lhs, rhs, z


def verify_step(hasher, step: StepProof, segment_feature_mask: int = 0) -> R.FsChallenges:
    """Full verification of one step proof; raises InvalidInput naming the first check that fails."""
    # the order of winter-verifier: the OOD frame is checked against the constraints BEFORE it is absorbed by the coin
    lhs, rhs, z = ood_consistency(hasher, step, air_of(step, segment_feature_mask))
    if lhs != rhs:
        raise InvalidInput("out-of-domain constraint evaluations are inconsistent with the composition polynomial")
    fs = R.replay_fs_from_step(hasher, step)
    if z != fs.z_ood:
        raise InvalidInput("out-of-domain point of the constraint check differs from the replayed transcript")
    R.verify_child_openings(hasher, step, fs)
    R.verify_child_deep_and_fri(step, fs)
    return fs
