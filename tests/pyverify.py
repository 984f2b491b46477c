"""Out-of-domain consistency check of a segment proof on Python integers, independent of oracle/: the constraints evaluated at z
from the OOD frame (tests/pyair.py), divided by their divisors and combined with the composition coefficients, must equal the
composition polynomial's columns recombined at z (winter-verifier `evaluate_constraints` + the check in `perform_verification`
[WF-mem]; periodic columns vm/air/mod.rs:520-592, vm/schedule.rs:35-80). Together with replay.py (transcript, openings, DEEP, FRI)
this is a second complete verifier next to the checker's C one."""
import pyair
from zk_lisp_b200 import replay as R
from zk_lisp_b200.field import P, inv, root_of_unity


def periodic_at(n: int, x: int):
    """31 cycle-32 columns (evaluated at x^(n/32)) and p_last (cycle n)."""
    y = pow(x, n // 32, P)
    w32 = root_of_unity(5)
    num = (pow(y, 32, P) - 1) * inv(32) % P
    L = [num * pow(w32, p, P) % P * inv((y - pow(w32, p, P)) % P) % P for p in range(32)]
    cols = [L[0]] + [L[1 + j] for j in range(27)] + [L[28], (L[29] + L[30] + L[31]) % P, L[31]]
    last = pow(root_of_unity(n.bit_length() - 1), n - 1, P)
    cols.append(last * inv(n) % P * (pow(x, n, P) - 1) % P * inv((x - last) % P) % P)
    return cols


def ood_check(zkl, hasher, step, air):
    """Returns (lhs, rhs); a valid proof has lhs == rhs."""
    v = step.proof.inner
    n, W = v.trace_len, v.main_width
    seed = zkl.air.context_to_elements(W, n, v.num_constraints, R._options_of(v)) + R.air_public_inputs_of(step).to_elements()
    coin = R.Coin(hasher, seed)
    trace_roots, constraint_root, _ = v.parse_commitments()
    coin.reseed(trace_roots[0])
    cc = coin.draw_many(v.num_constraints)           # transition coefficients, then one per assertion in sorted order
    coin.reseed(constraint_root)
    z = coin.draw()
    fe = lambda b, i: int.from_bytes(b[16 * i:16 * i + 16], "little")
    C = len(v.ood_constraints) // 32
    t_z, t_zg = [fe(v.ood_trace, c) for c in range(W)], [fe(v.ood_trace, W + c) for c in range(W)]
    h_z = [fe(v.ood_constraints, c) for c in range(C)]
    w0 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_0)
    w1 = zkl.params.rom_weights_for_seed(zkl.params.ROM_W_SEED_1)
    res = pyair.evaluate_transition(air, t_z, t_zg, periodic_at(n, z), w0, w1, zkl.field.be_from_le8(air.air_pi.core.merkle_root),
                                    rom_enabled=int(any(air.air_pi.core.program_commitment)))
    T = len(res)
    assert T + len(air.assertions) == v.num_constraints
    g = root_of_unity(n.bit_length() - 1)
    zn = pow(z, n, P)
    acc = sum(c * r for c, r in zip(cc, res)) % P
    lhs = acc * ((z - inv(g)) % P) % P * inv((zn - 1) % P) % P            # divisor (x^n - 1) / (x - g^(n-1))
    ordered = sorted(air.assertions, key=lambda t: (t[1], t[0]))          # by step, then column
    for k, (col, row, val) in enumerate(ordered):
        lhs = (lhs + cc[T + k] * (t_z[col] - val) % P * inv((z - pow(g, row, P)) % P)) % P
    rhs = sum(pow(zn, j, P) * h for j, h in enumerate(h_z)) % P
    return lhs, rhs, z
