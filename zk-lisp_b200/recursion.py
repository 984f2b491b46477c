"""Aggregation public inputs and the recursion artifact container `ZKLRC1` (second half).

Reference: zk-lisp-proof-winterfell/src/agg/pi.rs:20-218 (AggAirPublicInputs, profiles, to_elements),
src/lib.rs:395-480 (`RecursionPublicBuilder::build_public`: what a batch of step proofs contributes to the aggregation
statement), src/lib.rs:483-551 (encode) and :553-774 (decode). Host byte shuffling; the aggregation proof itself stays with
the reference's CPU prover and is carried here as opaque bytes.
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

from . import air as A
from .field import fold_bytes32_to_fe
from .step import InvalidInput, StepProof, children_root, public_inputs_digest

MAGIC = b"ZKLRC1"
_Z32 = bytes(32)


class RecursionInvalid(ValueError):
    """prove::Error::RecursionInvalid(msg)."""


@dataclass
class AggProfileMeta:      # agg/pi.rs:86-95
    m: int = 0
    rho: int = 0
    q: int = 0
    o: int = 0
    lambda_bits: int = 0
    pi_len: int = 0
    v_units: int = 0


@dataclass
class AggFriProfile:       # agg/pi.rs:99-105
    lde_blowup: int = 0
    folding_factor: int = 0
    redundancy: int = 0
    num_layers: int = 0


@dataclass
class AggQueryProfile:     # agg/pi.rs:109-113
    num_queries: int = 0
    grinding_factor: int = 0


@dataclass
class AggAirPublicInputs:
    """agg/pi.rs:20-82."""
    program_id: bytes = _Z32
    program_commitment: bytes = _Z32
    pi_digest: bytes = _Z32
    children_root: bytes = _Z32
    v_units_total: int = 0
    children_count: int = 0
    batch_id: bytes = _Z32
    profile_meta: AggProfileMeta = field(default_factory=AggProfileMeta)
    profile_fri: AggFriProfile = field(default_factory=AggFriProfile)
    profile_queries: AggQueryProfile = field(default_factory=AggQueryProfile)
    suite_id: bytes = _Z32
    children_ms: List[int] = field(default_factory=list)
    vm_state_initial: bytes = _Z32
    vm_state_final: bytes = _Z32
    ram_gp_unsorted_initial: bytes = _Z32
    ram_gp_unsorted_final: bytes = _Z32
    ram_gp_sorted_initial: bytes = _Z32
    ram_gp_sorted_final: bytes = _Z32
    rom_s_initial: Tuple[bytes, bytes, bytes] = (_Z32, _Z32, _Z32)
    rom_s_final: Tuple[bytes, bytes, bytes] = (_Z32, _Z32, _Z32)

    def to_elements(self) -> List[int]:
        """agg/pi.rs:176-217: the aggregation proof's coin seed tail."""
        f, pm, pf, pq = fold_bytes32_to_fe, self.profile_meta, self.profile_fri, self.profile_queries
        out = [f(self.program_id), f(self.program_commitment), f(self.pi_digest), f(self.children_root), f(self.batch_id)]
        out += [pm.m, pm.rho, pm.q, pm.o, pm.lambda_bits, pm.pi_len, pm.v_units, pf.lde_blowup, pf.folding_factor, pf.redundancy,
                pf.num_layers, pq.num_queries, pq.grinding_factor, self.children_count, self.v_units_total]
        out += [f(self.vm_state_initial), f(self.vm_state_final), f(self.ram_gp_unsorted_initial), f(self.ram_gp_unsorted_final),
                f(self.ram_gp_sorted_initial), f(self.ram_gp_sorted_final)]
        return out + [f(x) for x in self.rom_s_initial] + [f(x) for x in self.rom_s_final]


def _meta_of(step: StepProof) -> AggProfileMeta:
    m = step.proof.meta
    return AggProfileMeta(m.m, m.rho, m.q, m.o, m.lambda_bits, m.pi_len, m.v_units)


def from_step_proof(step: StepProof) -> AggAirPublicInputs:
    """agg/pi.rs:118-172: the statement of a batch holding this one child."""
    pi, meta = step.proof.pi, _meta_of(step)
    return AggAirPublicInputs(
        program_id=bytes(step.pi_core.program_id), program_commitment=bytes(step.pi_core.program_commitment),
        pi_digest=public_inputs_digest(step.pi_core),
        children_root=children_root(step.proof.header.suite_id, [(step.digest(), step.proof.root_trace)]),
        v_units_total=meta.v_units, children_count=1, profile_meta=meta,
        profile_fri=AggFriProfile(meta.rho, 2, 1, len(step.proof.inner.fri_layers)),
        profile_queries=AggQueryProfile(meta.q, 0), suite_id=step.proof.header.suite_id, children_ms=[meta.m],
        vm_state_initial=pi.state_in_hash, vm_state_final=pi.state_out_hash,
        ram_gp_unsorted_initial=pi.ram_gp_unsorted_in, ram_gp_unsorted_final=pi.ram_gp_unsorted_out,
        ram_gp_sorted_initial=pi.ram_gp_sorted_in, ram_gp_sorted_final=pi.ram_gp_sorted_out,
        rom_s_initial=(pi.rom_s_in_0, pi.rom_s_in_1, pi.rom_s_in_2), rom_s_final=(pi.rom_s_out_0, pi.rom_s_out_1, pi.rom_s_out_2))


def build_recursion_public(pub_inputs: A.PublicInputs, steps: Sequence[StepProof]) -> AggAirPublicInputs:
    """lib.rs:395-480: program binding from `pub_inputs`, children root over (step digest, root_trace) of every step, profile of
    the FIRST child, VM / RAM / ROM state of the first child's entry and the last child's exit."""
    if not steps:
        raise RecursionInvalid("build_recursion_public requires at least one step")
    for s in steps:
        c = s.pi_core
        if (bytes(c.program_id) != bytes(pub_inputs.program_id) or bytes(c.program_commitment) != bytes(pub_inputs.program_commitment)
                or list(c.main_args) != list(pub_inputs.main_args)):
            raise RecursionInvalid("build_recursion_public requires all steps to share the same program_id, program_commitment "
                                   "and main_args as recursion public inputs")
    first, last = from_step_proof(steps[0]), from_step_proof(steps[-1])
    suite_id = steps[0].proof.header.suite_id
    meta = first.profile_meta
    return AggAirPublicInputs(
        program_id=bytes(pub_inputs.program_id), program_commitment=bytes(pub_inputs.program_commitment),
        pi_digest=public_inputs_digest(pub_inputs),
        children_root=children_root(suite_id, [(s.digest(), s.proof.root_trace) for s in steps]),
        v_units_total=min(sum(s.proof.meta.v_units for s in steps), (1 << 64) - 1), children_count=len(steps),
        profile_meta=meta, profile_fri=AggFriProfile(meta.rho, 2, 1, 1), profile_queries=AggQueryProfile(meta.q, 0),
        suite_id=suite_id, children_ms=[s.proof.meta.m for s in steps],
        vm_state_initial=steps[0].state_in_hash(), vm_state_final=steps[-1].state_out_hash(),
        ram_gp_unsorted_initial=first.ram_gp_unsorted_initial, ram_gp_unsorted_final=last.ram_gp_unsorted_final,
        ram_gp_sorted_initial=first.ram_gp_sorted_initial, ram_gp_sorted_final=last.ram_gp_sorted_final,
        rom_s_initial=first.rom_s_initial, rom_s_final=last.rom_s_final)


def encode(proof_bytes: bytes, pi: AggAirPublicInputs) -> bytes:
    """lib.rs:486-551 (`proof_bytes` = winterfell::Proof::to_bytes of the aggregation proof)."""
    le = lambda v, n: int(v).to_bytes(n, "little")
    pm, pf, pq = pi.profile_meta, pi.profile_fri, pi.profile_queries
    out = bytearray(MAGIC)
    out += pi.program_id + pi.program_commitment + pi.pi_digest + pi.children_root + pi.batch_id
    out += le(pi.v_units_total, 8) + le(pi.children_count, 4)
    out += le(pm.m, 4) + le(pm.rho, 2) + le(pm.q, 2) + le(pm.o, 2) + le(pm.lambda_bits, 2) + le(pm.pi_len, 4) + le(pm.v_units, 8)
    out += le(pf.lde_blowup, 4) + le(pf.folding_factor, 1) + le(pf.redundancy, 1) + le(pf.num_layers, 1)
    out += le(pq.num_queries, 2) + le(pq.grinding_factor, 4)
    out += pi.suite_id + le(len(pi.children_ms), 4) + b"".join(le(m, 4) for m in pi.children_ms)
    out += (pi.vm_state_initial + pi.vm_state_final + pi.ram_gp_unsorted_initial + pi.ram_gp_unsorted_final
            + pi.ram_gp_sorted_initial + pi.ram_gp_sorted_final)
    out += b"".join(pi.rom_s_initial) + b"".join(pi.rom_s_final)
    out += le(len(proof_bytes), 4) + bytes(proof_bytes)
    return bytes(out)


def decode(data: bytes) -> Tuple[bytes, AggAirPublicInputs]:
    """lib.rs:553-774, same checks and messages; returns (aggregation proof bytes, public inputs)."""
    data = bytes(data)
    if len(data) < len(MAGIC) or data[:len(MAGIC)] != MAGIC:
        raise InvalidInput("invalid recursion artifact magic")
    i = [len(MAGIC)]

    def take(n, msg):
        if len(data) < i[0] + n:
            raise InvalidInput(msg)
        out = data[i[0]:i[0] + n]
        i[0] += n
        return out

    num = lambda n, msg: int.from_bytes(take(n, msg), "little")
    take32 = lambda name: take(32, name + " truncated")
    pi = AggAirPublicInputs()
    pi.program_id, pi.program_commitment, pi.pi_digest = take32("program_id"), take32("program_commitment"), take32("pi_digest")
    pi.children_root, pi.batch_id = take32("children_root"), take32("batch_id")
    pi.v_units_total, pi.children_count = num(8, "v_units_total truncated"), num(4, "children_count truncated")
    u32, u16, u8 = (lambda: num(4, "u32 truncated")), (lambda: num(2, "u16 truncated")), (lambda: num(1, "u8 truncated"))
    m, rho, q, o, lam, pi_len = u32(), u16(), u16(), u16(), u16(), u32()
    pi.profile_meta = AggProfileMeta(m, rho, q, o, lam, pi_len, num(8, "v_units(meta) truncated"))
    pi.profile_fri = AggFriProfile(u32(), u8(), u8(), u8())
    pi.profile_queries = AggQueryProfile(u16(), u32())
    pi.suite_id = take32("suite_id")
    pi.children_ms = [u32() for _ in range(u32())]
    pi.vm_state_initial, pi.vm_state_final = take32("vm_state_initial"), take32("vm_state_final")
    pi.ram_gp_unsorted_initial, pi.ram_gp_unsorted_final = take32("ram_gp_unsorted_initial"), take32("ram_gp_unsorted_final")
    pi.ram_gp_sorted_initial, pi.ram_gp_sorted_final = take32("ram_gp_sorted_initial"), take32("ram_gp_sorted_final")
    pi.rom_s_initial = tuple(take32("rom_s_initial lane") for _ in range(3))
    pi.rom_s_final = tuple(take32("rom_s_final lane") for _ in range(3))
    proof = take(num(4, "proof length truncated"), "proof bytes truncated")
    return proof, pi
