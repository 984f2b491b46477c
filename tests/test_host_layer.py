"""CPU tests of the host layer: the C-ABI library loads and exports every declared symbol, and the
product path refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as graft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(zkl):
    graft.build()
    header = open(os.path.join(ROOT, "include", "zkl_b200.h")).read()
    declared = set(re.findall(r"\b(zkl_[a-z0-9_]+)\s*\(", header))
    declared -= {"zkl_batch_proof"}
    lib = ctypes.CDLL(zkl.capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert declared == set(zkl.capi.EXPORTED_SYMBOLS)


def test_no_cpu_fallback(zkl):
    if _has_cuda():
        pytest.skip("CUDA device present")
    with pytest.raises(zkl.capi.ZklError) as e:
        zkl.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product package must never reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "zk-lisp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/zkl_b200.h must compile as C99 (what cgo / bindgen / ctypes-style binders read)
    and as C++, with nothing but plain pointers and sizes in the signatures."""
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "zkl_b200.h"\nint main(void) { zkl_child_challenges c; (void)c; return (int)sizeof(zkl_fe) - 16; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    header = open(os.path.join(inc, "zkl_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # comments may mention callers; declarations may not
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code


# ---- ZkLispAir::new in native code (zkl_program_derive / zkl_air_build) against the Python-integer mirror (air.py, params.py) ----
def _air_cases(zkl):
    import synth
    yield "valid-204", synth.valid_segment(n=64, seed=5)
    yield "valid-204-later-segment", synth.valid_segment(n=128, seed=6, start_level=12)
    yield "valid-full-layout", synth.valid_segment(n=64, seed=7, full_layout=True)
    yield "random-all", synth.random_segment(n=64, seed=8)
    yield "random-no-ram", synth.random_segment(n=64, seed=9, ram=False, usage_mask=0x0FF)
    yield "random-no-merkle-args", synth.random_segment(n=128, seed=10, merkle=False, usage_mask=0x155, ram_bits=0x8001,
                                                         main_args=[("u64", 7), ("bytes32", bytes(range(32))), ("u128", (1 << 127) + 5)], vm_expect=True)
    yield "random-pc-init", synth.random_segment(n=64, seed=11, pc_init=3, main_args=[("u64", 9)], usage_mask=0)
    yield "random-no-poseidon", synth.random_segment(n=64, seed=12, poseidon=False, sponge=False, usage_mask=0x17F)


def test_native_air_build_matches_python_mirror(zkl):
    """zkl_air_build must give exactly what air.py derives: column map, flags, degrees-derived counts, every assertion in
    get_assertions() order, every constant, and the public-coin seed elements."""
    from zk_lisp_b200 import prover as PV
    graft.build()
    for name, (trace, air_pi, opts) in _air_cases(zkl):
        W, n = trace.shape[0], trace.shape[1]
        a = zkl.air.Air(W, n, air_pi, opts)
        ref = a.to_c()
        nat = PV.ZkProver(None, opts, air_pi).native_air(W, n)
        got = nat.params
        for f in zkl.capi.COLUMN_FIELDS:
            assert getattr(got.cols, f) == getattr(ref.cols, f), (name, f)
        for f in ("f_poseidon", "f_vm", "f_sponge", "f_merkle", "f_ram", "rom_enabled", "vm_usage_mask", "ram_delta_clk_bits",
                  "num_transition", "ce_blowup", "num_composition_cols", "num_assertions"):
            assert getattr(got, f) == getattr(ref, f), (name, f)
        assert nat.num_assertions == len(a.assertions)
        for i, (col, step, val) in enumerate(a.assertions):
            x = nat.assertions[i]
            assert (x.column, x.step, x.value.int()) == (col, step, val), (name, i)
        assert bytes(got.pose) == bytes(ref.pose) and bytes(got.rom) == bytes(ref.rom), name
        assert bytes(got.rom_w_enc0) == bytes(ref.rom_w_enc0) and bytes(got.rom_w_enc1) == bytes(ref.rom_w_enc1), name
        assert got.program_fe0.int() == ref.program_fe0.int() and got.merkle_root_fe.int() == ref.merkle_root_fe.int(), name
        assert PV.fe_to_ints(nat.seed) == [v % zkl.field.P for v in a.seed_elements()], name


def test_closed_form_assertion_ranks_model(zkl):
    """The arithmetic behind csrc/api_stark.inl: upload_air_sched + k_expand_assertions on Python integers: from ONE level's template
    and the extras, template entry j of level l gets rank l T + j + (extras ordered before it), a template step's group starts at
    l T + first_j + (extras at earlier steps), extras and extra-only steps are ranked on the host — and the result is the explicit
    list sorted by (step, column) with one group per distinct step. The explicit list comes from the native `zkl_air_build` (every
    kind of extra in `_air_cases`: vm_expect on position 29, main_args, pc_init, ROM state at both ends). The GPU suite checks the
    kernel itself (`test_closed_form_assertions_equal_the_explicit_list`)."""
    from zk_lisp_b200 import prover as PV
    graft.build()
    for name, (trace, air_pi, opts) in _air_cases(zkl):
        W, n = trace.shape[0], trace.shape[1]
        nat = PV.ZkProver(None, opts, air_pi).native_air(W, n)
        explicit = [(nat.assertions[i].column, nat.assertions[i].step) for i in range(nat.num_assertions)]
        levels, period = n // 32, 32
        # the closed form: cells asserted at the same position of EVERY level are the template, the rest are extras
        per_pos = {}
        for col, step in explicit:
            per_pos.setdefault((step % period, col), set()).add(step // period)
        tmpl = sorted(k for k, lv in per_pos.items() if len(lv) == levels)                    # (pos, col), sorted
        tset = set(tmpl)
        extras = sorted((step, col) for col, step in explicit if (step % period, col) not in tset)   # (step, col), sorted
        T, E = len(tmpl), len(extras)
        assert levels * T + E == len(explicit), name
        tg = [(pos, j) for j, (pos, col) in enumerate(tmpl) if j == 0 or tmpl[j - 1][0] != pos]   # distinct positions, first_j
        G = len(tg)
        tpos = {pos for pos, _ in tg}
        xo = sorted({step for step, _ in extras if step % period not in tpos})                # steps only extras name
        ranked = [None] * len(explicit)
        ng = levels * G + len(xo)
        gs = [None] * (ng + 1)
        for l in range(levels):                                                               # k_expand_assertions
            for j, (pos, col) in enumerate(tmpl):
                step = l * period + pos
                eb = sum(1 for es, ec in extras if es < step or (es == step and ec < col))
                assert ranked[l * T + j + eb] is None
                ranked[l * T + j + eb] = (step, col)
            for k, (pos, first_j) in enumerate(tg):
                step = l * period + pos
                gs[l * G + k + sum(1 for x in xo if x < step)] = l * T + first_j + sum(1 for es, _ in extras if es < step)
        for e, (step, col) in enumerate(extras):                                              # upload_air_sched, host side
            before = sum(1 for pos, c in tmpl if pos < step % period or (pos == step % period and c < col))
            rank = (step // period) * T + before + e
            assert ranked[rank] is None
            ranked[rank] = (step, col)
        for x, step in enumerate(xo):
            lvl, pos = step // period, step % period
            gs[lvl * G + sum(1 for p_, _ in tg if p_ < pos) + x] = lvl * T + sum(1 for p_, _ in tmpl if p_ < pos) + sum(1 for es, _ in extras if es < step)
        gs[ng] = len(explicit)
        want = sorted((step, col) for col, step in explicit)
        assert ranked == want, name
        starts = [i for i in range(len(want)) if i == 0 or want[i][0] != want[i - 1][0]] + [len(want)]
        assert gs == starts, name


def test_native_air_build_rejects_what_the_mirror_rejects(zkl):
    import synth
    from zk_lisp_b200 import prover as PV
    trace, air_pi, opts = synth.valid_segment(n=64, seed=5)
    with pytest.raises(ValueError, match="does not match layout width"):
        PV.ZkProver(None, opts, air_pi).native_air(trace.shape[0] - 1, 64)
    with pytest.raises(ValueError, match="power of two"):
        PV.ZkProver(None, opts, air_pi).native_air(trace.shape[0], 48)
    air_pi.core.program_commitment = bytes(32)
    with pytest.raises(ValueError, match="both zero or both non-zero"):
        PV.ZkProver(None, opts, air_pi).native_air(trace.shape[0], 64)
    # expected-output register on the row / register that main_args also bind: the duplicate is dropped, first kept
    # (mod.rs:452-475), and the count then disagrees with what ZkLispAir::new declared: Winterfell panics, both mirrors refuse
    t, pi, o = synth.random_segment(n=64, seed=13, main_args=[("u64", 1), ("u64", 2)], vm_expect=True)
    pi.core.vm_out_reg, pi.core.vm_out_row = zkl.air.NR - 1, 0
    with pytest.raises(ValueError, match="declared 293 assertions but get_assertions yields 292"):
        zkl.air.Air(t.shape[0], 64, pi, o)
    with pytest.raises(ValueError, match="declared 293 assertions but get_assertions yields 292"):
        PV.ZkProver(None, o, pi).native_air(t.shape[0], 64)


def test_native_suite_derivation_matches_python(zkl):
    """zkl_suite_derive / zkl_program_derive (blake3-derived constants, Cauchy MDS, ROM suite, weights, field commitments)."""
    from zk_lisp_b200 import params, prover as PV
    from zk_lisp_b200.blake3 import blake3
    s, ref = PV.hasher_suite_c(), params.get_poseidon_suite(params.HASHER_SUITE_ID)
    assert bytes(s) == bytes(zkl.capi.suite_to_c(ref))
    for tag in (b"a", b"program-2"):
        pid, pc = blake3(tag), blake3(tag + b"/commitment")
        prog = PV.program_constants(pid, pc, cache=False)
        assert bytes(prog.pose) == bytes(zkl.capi.suite_to_c(params.get_poseidon_suite(pid)))
        assert bytes(prog.rom) == bytes(zkl.capi.rom_suite_to_c(*params.get_rom_suite(pid)))
        assert [prog.program_fe[i].int() for i in range(2)] == params.program_field_commitment(pid)
        assert [prog.commitment_fe[i].int() for i in range(2)] == params.program_field_commitment(pc)
        assert [prog.rom_w_enc0[k].int() for k in range(59)] == params.rom_weights_for_seed(params.ROM_W_SEED_0)
        assert [prog.rom_w_enc1[k].int() for k in range(59)] == params.rom_weights_for_seed(params.ROM_W_SEED_1)
    zero = PV.program_constants(bytes(32), bytes(32), cache=False)
    assert all(zero.program_fe[i].int() == 0 and zero.commitment_fe[i].int() == 0 for i in range(2))
