import random

import numpy as np

P = 2**128 - 45 * 2**40 + 1


def rand_fe(rng: random.Random, n: int):
    edge = [0, 1, 2, P - 1, P - 2, (1 << 64) - 1, 1 << 64, 1 << 127, (45 << 40) - 1, 45 << 40]
    return [rng.choice(edge) if rng.random() < 0.1 else rng.randrange(P) for _ in range(n)]


def to_fe(xs) -> np.ndarray:
    out = np.empty((len(xs), 2), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i, 0] = x & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = x >> 64
    return out


def from_fe(a: np.ndarray):
    a = np.asarray(a).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


def splitmix_trace(width: int, n: int, seed: int) -> np.ndarray:
    """SURVEY.md §8(d) S-synth generator: uniform field elements from splitmix64, two draws per
    element, rejection >= p. Returned column-major (W, n, 2) uint64."""
    mask = (1 << 64) - 1
    state = seed & mask
    out = np.empty((width, n, 2), dtype=np.uint64)

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return z ^ (z >> 31)

    for c in range(width):
        for r in range(n):
            while True:
                lo, hi = nxt(), nxt()
                if (hi << 64 | lo) < P:
                    break
            out[c, r, 0], out[c, r, 1] = lo, hi
    return out


def fast_random_trace(width: int, n: int, seed: int) -> np.ndarray:
    """numpy-speed random canonical elements (hi limb < 2^64 - 2^32 keeps the value below p)."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 2**63, size=(width, n, 2), dtype=np.uint64) * np.uint64(2)
    out[..., 1] = out[..., 1] >> np.uint64(1)  # < 2^63 < p's high limb
    out[..., 0] |= rng.integers(0, 2, size=(width, n), dtype=np.uint64)
    return out


def single_segment_step(zkl, air_pi, opts, n, proof, min_bits=64):
    """zl1 step container of a single-segment proof, boundary values taken from the AIR public inputs
    (what prove.rs:1108-1172 assembles around the inner proof)."""
    S = zkl.step
    f = S.fe_to_bytes_fold
    meta = S.StepMeta.from_env(n, opts.blowup, opts.num_queries, min_bits, len(air_pi.to_elements()))
    b = {"ram_gp_unsorted_in": f(air_pi.ram_gp_unsorted_in), "ram_gp_unsorted_out": f(air_pi.ram_gp_unsorted_out),
         "ram_gp_sorted_in": f(air_pi.ram_gp_sorted_in), "ram_gp_sorted_out": f(air_pi.ram_gp_sorted_out),
         "state_in_hash": bytes(32), "state_out_hash": bytes(32)}
    for i in range(3):
        b[f"rom_s_in_{i}"], b[f"rom_s_out_{i}"] = f(air_pi.rom_s_in[i]), f(air_pi.rom_s_out[i])
    zl1 = S.Zl1Proof.new_single_segment(air_pi.core.program_id, meta, air_pi.core, f(air_pi.pc_init), b, proof)
    return S.StepProof(zl1, air_pi.core, tuple(air_pi.rom_acc), air_pi.vm_usage_mask, air_pi.ram_delta_clk_bits)
