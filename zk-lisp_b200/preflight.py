"""Preflight of a trace against the AIR before proving — the optional first step of `ZkProver::prove`
(zk-lisp-proof-winterfell/src/prove.rs:186-211 -> src/preflight.rs:71-420; modes zk-lisp-proof/src/frontend.rs:23-27).

Every row but the exempt last one is evaluated against `evaluate_transition` with the periodic columns of that row; the first
non-zero constraint aborts the proof with the reference's message and a report of what the row looked like. It is a debugging aid
(off by default) that walks the host trace on Python integers (air_eval.py): ~1 ms per row, never on the proving path proper.
"""
import json
from dataclasses import asdict, dataclass, field
from typing import List, Optional, Tuple

from . import air as A
from . import air_eval, params
from .field import P, be_from_le8, root_of_unity
from .verify import periodic_at

OFF, CONSOLE, JSON = "off", "console", "json"


class PreflightError(RuntimeError):
    """prove::Error::Backend("preflight: constraint {i} non-zero at row {r}: {v:?}")."""

    def __init__(self, report):
        super().__init__(f"preflight: constraint {report.constraint_idx} non-zero at row {report.row}: {report.constraint_val}")
        self.report = report


@dataclass
class PreflightReport:
    """preflight.rs:29-45 (the gadget snapshots the console renderer prints are the lanes and the schedule gates)."""
    row: int
    constraint_idx: int
    constraint_val: int
    pos: int
    gates_map: bool
    gates_final: bool
    rounds: List[bool]
    lanes_cur: Tuple[int, int, int, int]
    lanes_next: Tuple[int, int, int, int]
    lanes_exp: Optional[Tuple[int, int, int, int]] = None
    block: str = field(default="")


def _block_of(air: A.Air, idx: int, sizes) -> str:
    for name, size in sizes:
        if idx < size:
            return name
        idx -= size
    return "?"


def run(mode: str, air: A.Air, trace) -> Optional[PreflightReport]:
    """preflight.rs:71-420. `trace`: (W, n, 2) uint64 host array. Returns None when every row passes; raises PreflightError
    (after printing the report in console / json mode) on the first failing constraint."""
    if mode == OFF:
        return None
    W, n = trace.shape[0], trace.shape[1]
    if W != air.width or n != air.n:
        raise ValueError("preflight: trace shape does not match the AIR")
    rows = [[int(trace[c, r, 0]) | (int(trace[c, r, 1]) << 64) for c in range(W)] for r in range(n)]
    core = air.air_pi.core
    w0, w1 = params.rom_weights_for_seed(params.ROM_W_SEED_0), params.rom_weights_for_seed(params.ROM_W_SEED_1)
    root_fe, rom = be_from_le8(core.merkle_root), int(any(core.program_commitment))
    g = root_of_unity(n.bit_length() - 1)
    x = 1
    cols = air.cols
    for r in range(n - 1):                      # the last row is the transition exemption
        per = periodic_at(n, x)
        res = air_eval.evaluate_transition(air, rows[r], rows[r + 1], per, w0, w1, root_fe, rom_enabled=rom)
        bad = next(((i, v) for i, v in enumerate(res) if v), None)
        if bad is not None:
            i, v = bad
            cur, nxt = rows[r], rows[r + 1]
            lanes = lambda row: (row[cols.lane_l], row[cols.lane_r], row[cols.lane_c0], row[cols.lane_c1])
            exp = None
            if air.fm & A.FM_POSEIDON and i < 12 * A.POSEIDON_ROUNDS:      # expected next lanes of the failing round (preflight.rs:204-240)
                j = i // 12
                s3 = [pow(cur[cols.lane_index(k)], 3, P) for k in range(12)]
                y = [(sum(air.pose.mds[a][k] * s3[k] for k in range(12)) + air.pose.rc[j][a]) % P for a in range(12)]
                exp = (y[cols.lane_l], y[cols.lane_r], y[cols.lane_c0], y[cols.lane_c1])
            sizes = []
            if air.fm & A.FM_POSEIDON:
                sizes.append(("poseidon", len(air_eval.poseidon_block(air, cur, nxt, per))))
            if air.fm & A.FM_VM:
                sizes += [("vm-ctrl", len(air_eval.ctrl_block(air, cur, nxt, per, rom))), ("vm-alu", len(air_eval.alu_block(air, cur, nxt, per)))]
            if air.fm & A.FM_RAM:
                sizes.append(("ram", len(air_eval.ram_block(air, cur, nxt, per))))
            if air.fm & A.FM_MERKLE:
                sizes.append(("merkle", 7))
            sizes.append(("rom", 86))
            report = PreflightReport(r, i, v, r % A.STEPS_PER_LEVEL_P2, bool(cur[cols.g_map]), bool(cur[cols.g_final]),
                                     [bool(cur[cols.g_r_index(k)]) for k in range(A.POSEIDON_ROUNDS)], lanes(cur), lanes(nxt), exp,
                                     _block_of(air, i, sizes))
            if mode == JSON:
                def js(x_):          # field elements do not fit JSON numbers: decimal strings, like the reference's report
                    if isinstance(x_, (list, tuple)):
                        return [js(y_) for y_ in x_]
                    return str(x_) if isinstance(x_, int) and not isinstance(x_, bool) and x_ >= 1 << 53 else x_
                print(json.dumps({k: js(x_) for k, x_ in asdict(report).items()}, indent=1))
            elif mode == CONSOLE:
                print(f"preflight: FAILED row {r} (level {r // 32}, pos {report.pos}) constraint {i} [{report.block}] = {v}")
                print(f"  gates: map={report.gates_map} final={report.gates_final} round={[k for k, b in enumerate(report.rounds) if b]}")
                print(f"  lanes cur  {report.lanes_cur}\n  lanes next {report.lanes_next}" + (f"\n  lanes exp  {exp}" if exp else ""))
            raise PreflightError(report)
        x = x * g % P
    if mode == CONSOLE:
        print(f"preflight: OK (rows={n}, constraints={len(air.degrees)})")
    return None
