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


HASHER_COLLISION_RESISTANCE = 128   # poseidon/hasher.rs:60
GRINDING_CONTRIBUTION_FLOOR = 80    # prove.rs:1185


def conjectured_security_bits(num_queries: int, blowup: int, grinding: int, extension_degree: int, lde_domain_size: int) -> int:
    """Conjectured security of a set of proof options, the quantity `AcceptableOptions::MinConjecturedSecurity(min_bits)` bounds
    when the reference verifies (prove.rs:738, 814). zk-lisp restates the estimate at prove.rs:1177-1195 (128 field bits per
    extension degree, log2(blowup) bits per query, grinding counted only from 80 query bits up, minus one, capped by the hasher's
    collision resistance); winter-air additionally charges the field for the LDE domain size (field bits - log2(N)) [WF-mem].
    The stricter of the two is returned."""
    if blowup < 2 or blowup & (blowup - 1) or lde_domain_size < 1:
        return 0
    field_security = 128 * extension_degree - (lde_domain_size.bit_length() - 1)
    query_security = (blowup.bit_length() - 1) * num_queries
    if query_security >= GRINDING_CONTRIBUTION_FLOOR:
        query_security += grinding
    return max(min(min(field_security, query_security) - 1, HASHER_COLLISION_RESISTANCE), 0)


def check_acceptable(step: StepProof, min_bits: int) -> int:
    """What the stock verifier settles before any cryptography: the proof's self-declared options must reach `min_bits` of
    conjectured security (a forged proof declaring q = 1, grind = 0, blowup = 2 replays consistently otherwise), must be the
    options zk-lisp proves steps with (FieldExtension::None, prove.rs:967; Linear batching, :970-971; folding factor 2, :968) and
    every field element of the proof must be canonical (winter-math's reader rejects values >= p). Returns the security level."""
    v = step.proof.inner
    if v.num_queries == 0 or v.num_unique_queries == 0:
        raise InvalidInput("proof declares no queries")
    if v.field_extension != 1:
        raise InvalidInput("step proofs use FieldExtension::None")
    if tuple(v.batching) != (0, 0):
        raise InvalidInput("step proofs use Linear batching")
    if v.fri_folding != 2:
        raise InvalidInput("step proofs use FRI folding factor 2")
    if v.blowup < 2 or v.blowup & (v.blowup - 1):
        raise InvalidInput("blowup factor must be a power of two >= 2")
    bits = conjectured_security_bits(v.num_queries, v.blowup, v.grinding, 1, v.lde_domain_size)
    if bits < min_bits:
        raise InvalidInput(f"proof options give {bits} bits of conjectured security, below the required {min_bits}")
    blobs = [v.trace_query_values, v.constraint_query_values, v.ood_trace, v.ood_constraints, v.fri_remainder] + [vals for vals, _ in v.fri_layers]
    for blob in blobs:
        if len(blob) % 16:
            raise InvalidInput("field element section of the proof is not a multiple of 16 bytes")
        for i in range(0, len(blob), 16):
            if int.from_bytes(blob[i:i + 16], "little") >= P:
                raise InvalidInput("non-canonical field element in the proof")
    return bits


def verify_step(hasher, step: StepProof, min_bits: int, segment_feature_mask: int = 0) -> R.FsChallenges:
    """Full verification of one step proof; raises InvalidInput naming the first check that fails. `min_bits` is mandatory, like
    the reference's `min_security_bits` (zk-lisp-proof/src/lib.rs:44,59: 128 in release builds, 64 in debug builds)."""
    check_acceptable(step, min_bits)
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
