"""AIR-valid traces with every gadget live (tests/synth.py: rich_segment): RAM with the sorted table and the delta_clk range gadget,
eq / select / assert / assert_bit, 32- and 64-bit range checks, divmod / mulwide / div128, sponge absorb + squeeze, a Merkle path.
CPU side: the traces satisfy the Python-integer AIR row by row, the checker proves and verifies one, and the full-size golden
fixture (tests/golden/segment_golden.json) still belongs to the traces the generators produce."""
import hashlib
import json
import os

import pytest

import synth
from util import from_fe

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segment_golden.json")


def _rows(trace):
    W, n = trace.shape[0], trace.shape[1]
    return [[int(trace[c, r, 0]) | (int(trace[c, r, 1]) << 64) for c in range(W)] for r in range(n)]


@pytest.mark.parametrize("merkle", [False, True])
def test_rich_segment_satisfies_python_air(zkl, merkle):
    from zk_lisp_b200 import air_eval, params
    from zk_lisp_b200.field import P, be_from_le8, root_of_unity
    n = 1024
    trace, air_pi, opts = synth.rich_segment(n=n, seed=5, merkle=merkle)
    assert trace.shape[0] == (219 if merkle else 212)
    assert air_pi.vm_usage_mask == 0x1FF and air_pi.ram_delta_clk_bits != 0     # every optional constraint set is switched on
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    rows = _rows(trace)
    w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
    root_fe = be_from_le8(air_pi.core.merkle_root)
    g, x = root_of_unity(n.bit_length() - 1), 1
    for r in range(n - 1):
        res = air_eval.evaluate_transition(a, rows[r], rows[r + 1], zkl.verify.periodic_at(n, x), w0, w1, root_fe)
        assert len(res) == len(a.degrees)
        assert not any(v % P for v in res), (r, [i for i, v in enumerate(res) if v % P][:8])
        x = x * g % P
    for col, row, val in a.assertions:
        assert rows[row][col] == val % P
    # the gadgets really are in the trace: loads, stores, sorted RAM rows, range bits, a squeeze, the Merkle gate
    c = a.cols
    col = lambda k: [rows[r][k] for r in range(n)]
    assert sum(col(c.op_load)) >= 4 and sum(col(c.op_store)) >= 8 and sum(col(c.ram_start)) >= 6
    assert sum(col(c.op_divmod)) == 2 and sum(col(c.op_div128)) == 2 and sum(col(c.op_mulwide)) == 2 and sum(col(c.op_assert_range)) == 6
    assert sum(col(c.op_sponge)) == 6 and sum(col(c.pose_active)) >= 32
    if merkle:
        assert sum(col(c.merkle_start)) == 4 * 32 and any(air_pi.core.merkle_root)


def test_checker_proves_and_verifies_a_rich_segment(zkl, oracle):
    n = 1024
    trace, air_pi, opts = synth.rich_segment(n=n, seed=6, queries=8, grind=4)
    a = zkl.air.Air(trace.shape[0], n, air_pi, opts)
    ac, oc, seed = a.to_c(), opts.to_c(), zkl.prover.ints_to_fe(a.seed_elements())
    proof = oracle.prove_segment(trace, ac, seed, oc)
    rc, msg = oracle.verify_segment(proof, trace.shape[0], n, ac, seed, oc)
    assert rc == 0, msg


def test_golden_fixture_belongs_to_the_generated_traces(zkl):
    """tests/golden/segment_golden.json holds the checker's proofs of full-size segments (generator: make_segment_golden.py). The
    GPU suite compares proof bytes against it; here, without a GPU: the traces the generators produce today are the traces the
    fixture was made from, and the AIR shape recorded with them is what both AIR builders derive."""
    from zk_lisp_b200 import prover as PV
    gold = json.load(open(GOLDEN))
    assert {"rollup-204", "rollup-212-ram", "full-219-merkle"} <= set(gold)
    for name, rec in gold.items():
        trace, air_pi, opts = getattr(synth, rec["gen"])(**rec["kw"])
        assert hashlib.blake2b(trace.tobytes(), digest_size=16).hexdigest() == rec["trace_blake2b"], name
        assert (trace.shape[0], trace.shape[1]) == (rec["width"], rec["n"])
        assert (air_pi.vm_usage_mask, air_pi.ram_delta_clk_bits) == (rec["vm_usage_mask"], rec["ram_delta_clk_bits"])
        nat = PV.ZkProver(None, opts, air_pi).native_air(rec["width"], rec["n"])
        assert (nat.num_transition, nat.num_assertions) == (rec["num_transition"], rec["num_assertions"]), name
