"""Step-proof container around a segment proof: the `zl1` format, the step digest and the `ZKLSTP1` codec
(SURVEY.md §8f N3). Host byte shuffling plus a handful of sponge permutations per proof, exactly as in the reference.

Reference: zk-lisp-proof-winterfell/src/proof/step.rs (StepMeta :22-40,498-534; StepProof :42-56; to_bytes :79-152;
from_bytes :154-495), src/proof/format.rs (Header :19-38, PublicInputs :40-83, new_single_segment :160-258 incl. the
blake3 `root_trace` over the inner commitments :219-243), src/proof/digest.rs:15-68 (step digest),
src/poseidon/mod.rs:255-291 (poseidon_hash_two_lanes), src/agg/child.rs:853-895 (children root),
src/utils.rs:312-339 (VM state hash of a row), zk-lisp-proof/src/pi.rs:113-147 (public-input digest).

The inner Winterfell proof is NOT re-encoded: `InnerProof` is a read-only view over the bytes `zkl_prove_segment`
returned (layout written in csrc/api_stark.inl, winter-air 0.13 `Proof::to_bytes` [WF-mem]).
"""
from dataclasses import dataclass, field
from functools import lru_cache
from typing import List, Sequence, Tuple

from . import air as A
from . import params
from .blake3 import blake3
from .field import P, fold_bytes32_to_fe

MAGIC = b"ZKLSTP1"
U64_MAX = (1 << 64) - 1


class InvalidInput(ValueError):
    """zk_lisp_proof::error::Error::InvalidInput(msg)."""


# ---------------------------------------------------------------------------------------------------------------
# small helpers (utils.rs)
# ---------------------------------------------------------------------------------------------------------------
def fe_to_bytes_fold(x: int) -> bytes:
    """utils.rs:375-381."""
    return (x % P).to_bytes(16, "little") + bytes(16)


def fe_from_bytes_fold(b32: bytes) -> int:
    """utils.rs:386-390 (low 16 bytes, reduced; the upper half is ignored)."""
    return int.from_bytes(b32[0:16], "little") % P


def poseidon_hash_two_lanes_py(suite_id: bytes, left: int, right: int) -> int:
    """poseidon/mod.rs:255-291 on Python integers: one permutation of [left, right, 0 x 8, dom0, dom1] under the suite of
    `suite_id` (second implementation, used by the tests to check the native one)."""
    suite = params.get_poseidon_suite(bytes(suite_id))
    state = [left % P, right % P] + [0] * 8 + list(suite.dom)
    return params.permute(suite, state)[0]


@lru_cache(maxsize=64)
def _suite_struct(suite_id: bytes):
    from . import capi
    return capi.suite_to_c(params.get_poseidon_suite(suite_id))


def hash_two_lanes_batch(suite_id: bytes, pairs: Sequence[Tuple[int, int]]) -> List[int]:
    """poseidon_hash_two_lanes for a batch of (left, right) in native code (`zkl_hash_two_lanes_host`: the kernels' limb
    arithmetic compiled for the host; no device involved)."""
    import ctypes as C
    from . import capi
    k = len(pairs)
    if k == 0:
        return []
    lib = capi.load_library()
    left = (capi.Fe * k)(*[capi.Fe.of(l % P) for l, _ in pairs])
    right = (capi.Fe * k)(*[capi.Fe.of(r % P) for _, r in pairs])
    out = (capi.Fe * k)()
    if lib.zkl_hash_two_lanes_host(C.byref(_suite_struct(bytes(suite_id))), left, right, k, out):
        raise InvalidInput("zkl_hash_two_lanes_host failed")
    return [o.int() for o in out]


def poseidon_hash_two_lanes(suite_id: bytes, left: int, right: int) -> int:
    """poseidon/mod.rs:255-291."""
    return hash_two_lanes_batch(suite_id, [(left, right)])[0]


def vm_state_hash(registers: Sequence[int]) -> bytes:
    """utils.rs:312-339: blake3("zkl/vm/state-v1" || r0..r7 as 16-byte LE) of one trace row."""
    assert len(registers) == A.NR
    return blake3(b"zkl/vm/state-v1" + b"".join((r % P).to_bytes(16, "little") for r in registers))


def encode_main_args(args) -> bytes:
    """tag + LE bytes per argument (step.rs:97-115, pi.rs:123-140)."""
    out = len(args).to_bytes(4, "little")
    for kind, v in args:
        if kind == "u64":
            out += b"\x00" + int(v).to_bytes(8, "little")
        elif kind == "u128":
            out += b"\x01" + int(v).to_bytes(16, "little")
        elif kind == "bytes32":
            assert len(v) == 32
            out += b"\x02" + bytes(v)
        else:
            raise InvalidInput("unknown VmArg kind")
    return out


def public_inputs_digest(pi: A.PublicInputs) -> bytes:
    """zk-lisp-proof/src/pi.rs:113-147."""
    return blake3(b"zkl/pi/v1" + bytes(pi.program_id) + bytes(pi.program_commitment) + bytes(pi.merkle_root)
                  + pi.feature_mask.to_bytes(8, "little") + encode_main_args(pi.main_args))


# ---------------------------------------------------------------------------------------------------------------
# read-only view of the inner Winterfell proof bytes
# ---------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, b: bytes):
        self.b, self.i = b, 0

    def take(self, n: int) -> bytes:
        if self.i + n > len(self.b):
            raise InvalidInput("failed to decode inner Winterfell proof in step proof")
        out = self.b[self.i:self.i + n]
        self.i += n
        return out

    def u8(self):
        return self.take(1)[0]

    def uint(self, n):
        return int.from_bytes(self.take(n), "little")

    def usize(self) -> int:
        """winter-utils vint64: the number of trailing zero bits of the first byte gives the length."""
        first = self.b[self.i] if self.i < len(self.b) else None
        if first is None:
            raise InvalidInput("failed to decode inner Winterfell proof in step proof")
        if first == 0:
            self.i += 1
            return self.uint(8)
        length = (first & -first).bit_length()
        return self.uint(length) >> length


@dataclass
class InnerProof:
    """Fields of `winterfell::Proof` the container and the aggregation layer read."""
    main_width: int
    aux_width: int
    trace_len: int
    num_constraints: int
    num_queries: int
    blowup: int
    grinding: int
    field_extension: int
    fri_folding: int
    fri_remainder_max_degree: int
    num_partitions: int
    hash_rate: int
    num_unique_queries: int
    commitments: bytes
    trace_query_values: bytes
    trace_query_paths: bytes
    constraint_query_values: bytes
    constraint_query_paths: bytes
    ood_trace: bytes
    ood_constraints: bytes
    fri_layers: List[Tuple[bytes, bytes]]
    fri_remainder: bytes
    fri_num_partitions: int
    pow_nonce: int
    batching: Tuple[int, int] = (0, 0)

    @property
    def lde_domain_size(self) -> int:
        return self.trace_len * self.blowup

    @property
    def num_fri_layers(self) -> int:
        """FriOptions::num_fri_layers(lde_domain_size) [WF-mem]."""
        domain, layers = self.lde_domain_size, 0
        max_rem = (self.fri_remainder_max_degree + 1) * self.blowup
        while domain > max_rem:
            domain //= self.fri_folding
            layers += 1
        return layers

    def parse_commitments(self) -> Tuple[List[bytes], bytes, List[bytes]]:
        """`Commitments::parse(num_trace_segments = 1, num_fri_layers)`: trace roots, constraint root, the FRI layer
        roots followed by the remainder commitment; trailing or missing bytes are an error."""
        want = 32 * (1 + 1 + self.num_fri_layers + 1)
        if len(self.commitments) != want:
            raise InvalidInput("invalid Winterfell commitments layout in zl1::Proof")
        d = [self.commitments[i:i + 32] for i in range(0, want, 32)]
        return d[:1], d[1], d[2:]

    def fri_remainder_coeffs(self) -> List[int]:
        """`FriProof::parse_remainder::<BE>()`."""
        if len(self.fri_remainder) % 16:
            raise InvalidInput("invalid FRI remainder layout in ZlChildTranscript")
        return [int.from_bytes(self.fri_remainder[i:i + 16], "little") for i in range(0, len(self.fri_remainder), 16)]

    def trace_rows(self) -> List[List[int]]:
        return _rows(self.trace_query_values, self.main_width, self.num_unique_queries)

    def constraint_rows(self) -> List[List[int]]:
        if self.num_unique_queries == 0:
            return []
        width = len(self.constraint_query_values) // (16 * self.num_unique_queries)
        return _rows(self.constraint_query_values, width, self.num_unique_queries)

    @staticmethod
    def parse(data: bytes) -> "InnerProof":
        r = _Reader(bytes(data))
        main_width, aux_width, log_n = r.u8(), r.u8(), r.u8()
        r.take(r.uint(2))                                   # trace meta
        if main_width == 0 or log_n >= 64:
            raise InvalidInput("failed to decode inner Winterfell proof in step proof")
        r.take(r.u8())                                      # field modulus bytes
        num_constraints = r.usize()
        q, blowup, grind, ext, folding, rem_deg = (r.u8() for _ in range(6))
        batching = tuple(r.take(2))                         # batching methods (constraints, DEEP); 0 = Linear
        num_partitions, hash_rate = r.u8(), r.u8()
        if blowup == 0 or folding < 2:
            raise InvalidInput("failed to decode inner Winterfell proof in step proof")
        unique = r.u8()
        commitments = r.take(r.uint(2))
        tq_values = r.take(r.usize())
        tq_paths = r.take(r.usize())
        cq_values = r.take(r.usize())
        cq_paths = r.take(r.usize())
        ood_trace = r.take(r.uint(2))
        ood_cons = r.take(r.uint(2))
        layers = []
        for _ in range(r.u8()):
            values = r.take(r.uint(4))
            layers.append((values, r.take(r.uint(4))))
        remainder = r.take(r.uint(2))
        fri_parts = r.u8()
        nonce = r.uint(8)
        if r.i != len(r.b):
            raise InvalidInput("failed to decode inner Winterfell proof in step proof")
        p = InnerProof(main_width, aux_width, 1 << log_n, num_constraints, q, blowup, grind, ext, folding, rem_deg,
                       num_partitions, hash_rate, unique, commitments, tq_values, tq_paths, cq_values, cq_paths,
                       ood_trace, ood_cons, layers, remainder, fri_parts, nonce)
        p.batching = batching
        return p


def _rows(values: bytes, width: int, count: int) -> List[List[int]]:
    if width == 0 or len(values) != 16 * width * count:
        raise InvalidInput("invalid trace queries layout in ZlChildTranscript")
    return [[int.from_bytes(values[16 * (q * width + c):16 * (q * width + c) + 16], "little") for c in range(width)]
            for q in range(count)]


# ---------------------------------------------------------------------------------------------------------------
# zl1 container (proof/format.rs) and step metadata (proof/step.rs)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class StepMeta:
    """step.rs:22-40. `o` counts the committed matrices (main trace + composition)."""
    m: int = 0
    rho: int = 0
    q: int = 0
    o: int = 0
    lambda_bits: int = 0
    pi_len: int = 0
    v_units: int = 0

    @staticmethod
    def new(m, rho, q, o, lambda_bits, pi_len) -> "StepMeta":
        """step.rs:499-512 (saturating m * q)."""
        return StepMeta(m, rho, q, o, lambda_bits, pi_len, min(m * q, U64_MAX))

    @staticmethod
    def from_env(trace_len: int, blowup: int, num_queries: int, lambda_bits: int, pi_len: int) -> "StepMeta":
        """step.rs:514-533."""
        return StepMeta.new(trace_len & 0xFFFFFFFF, blowup & 0xFFFF, num_queries & 0xFFFF, 2, min(lambda_bits, 0xFFFF), pi_len)

    def to_bytes(self) -> bytes:
        """digest.rs:21-28 field order and widths."""
        return (self.m.to_bytes(4, "little") + self.rho.to_bytes(2, "little") + self.q.to_bytes(2, "little")
                + self.o.to_bytes(2, "little") + self.lambda_bits.to_bytes(2, "little") + self.pi_len.to_bytes(4, "little")
                + self.v_units.to_bytes(8, "little"))


@dataclass
class Header:
    """format.rs:19-38."""
    suite_id: bytes
    rho: int
    q: int
    profile_id: int = 1
    field_id: int = 1
    version: int = 1
    fri_folding: int = 2
    ext: int = 0


BOUNDARY_FIELDS = ("state_in_hash", "state_out_hash", "ram_gp_unsorted_in", "ram_gp_unsorted_out", "ram_gp_sorted_in",
                   "ram_gp_sorted_out", "rom_s_in_0", "rom_s_in_1", "rom_s_in_2", "rom_s_out_0", "rom_s_out_1", "rom_s_out_2")
_Z32 = bytes(32)


@dataclass
class Zl1PublicInputs:
    """format.rs:40-83; every boundary value is a 32-byte folded field element or a blake3 state hash."""
    program_id: bytes = _Z32
    program_commitment: bytes = _Z32
    feature_mask: int = 0
    segment_index: int = 0
    segments_total: int = 1
    pc_init: bytes = _Z32
    state_in_hash: bytes = _Z32
    state_out_hash: bytes = _Z32
    ram_gp_unsorted_in: bytes = _Z32
    ram_gp_unsorted_out: bytes = _Z32
    ram_gp_sorted_in: bytes = _Z32
    ram_gp_sorted_out: bytes = _Z32
    rom_s_in_0: bytes = _Z32
    rom_s_in_1: bytes = _Z32
    rom_s_in_2: bytes = _Z32
    rom_s_out_0: bytes = _Z32
    rom_s_out_1: bytes = _Z32
    rom_s_out_2: bytes = _Z32

    def boundary_bytes(self) -> bytes:
        return b"".join(getattr(self, f) for f in BOUNDARY_FIELDS)


def root_trace(suite_id: bytes, inner: InnerProof) -> bytes:
    """format.rs:219-243: blake3("zkl/step/root_trace" || suite_id || trace roots || constraint root || FRI roots)."""
    trace_roots, constraint_root, fri_roots = inner.parse_commitments()
    return blake3(b"zkl/step/root_trace" + bytes(suite_id) + b"".join(trace_roots) + constraint_root + b"".join(fri_roots))


@dataclass
class Zl1Proof:
    """format.rs:93-100."""
    header: Header
    pi: Zl1PublicInputs
    root_trace: bytes
    meta: StepMeta
    inner_bytes: bytes
    inner: InnerProof

    @staticmethod
    def new_single_segment(suite_id: bytes, meta: StepMeta, core: A.PublicInputs, pc_init: bytes, boundaries: dict,
                           inner_bytes: bytes) -> "Zl1Proof":
        """format.rs:160-258. `boundaries` maps BOUNDARY_FIELDS to 32-byte values."""
        inner = InnerProof.parse(inner_bytes)
        header = Header(bytes(suite_id), meta.rho, meta.q)
        pi = Zl1PublicInputs(bytes(core.program_id), bytes(core.program_commitment), core.feature_mask, 0, 1, bytes(pc_init),
                             **{f: bytes(boundaries[f]) for f in BOUNDARY_FIELDS})
        return Zl1Proof(header, pi, root_trace(suite_id, inner), meta, bytes(inner_bytes), inner)

    @staticmethod
    def new_multi_segment(suite_id, meta, core, segment_index, segments_total, pc_init, boundaries, inner_bytes) -> "Zl1Proof":
        """format.rs:104-157."""
        p = Zl1Proof.new_single_segment(suite_id, meta, core, pc_init, boundaries, inner_bytes)
        p.pi.segment_index, p.pi.segments_total = segment_index, segments_total
        return p


def step_digest(proof: Zl1Proof) -> bytes:
    """digest.rs:15-68."""
    sid = proof.header.suite_id
    suite_fe = params.ro_from_slices("zkl/step/digest/suite", [sid])
    h_meta = poseidon_hash_two_lanes(sid, params.ro_from_slices("zkl/step/digest/meta", [proof.meta.to_bytes()]), 0)
    pi = proof.pi
    pi_bytes = (pi.program_id + pi.program_commitment + pi.feature_mask.to_bytes(8, "little")
                + pi.segment_index.to_bytes(4, "little") + pi.segments_total.to_bytes(4, "little") + pi.pc_init
                + pi.boundary_bytes())
    h_pi = poseidon_hash_two_lanes(sid, params.ro_from_slices("zkl/step/digest/pi", [pi_bytes]), 0)
    h_roots = poseidon_hash_two_lanes(sid, fold_bytes32_to_fe(proof.root_trace), 0)
    c0 = poseidon_hash_two_lanes(sid, suite_fe, h_meta)
    c1 = poseidon_hash_two_lanes(sid, c0, h_pi)
    return fe_to_bytes_fold(poseidon_hash_two_lanes(sid, c1, h_roots))


@dataclass
class StepProof:
    """step.rs:42-56."""
    proof: Zl1Proof
    pi_core: A.PublicInputs
    rom_acc: Tuple[int, int, int] = (0, 0, 0)
    vm_usage_mask: int = 0
    ram_delta_clk_bits: int = 0
    _digest: bytes = field(default=None, repr=False, compare=False)

    def digest(self) -> bytes:
        if self._digest is None:
            self._digest = step_digest(self.proof)
        return self._digest

    def state_in_hash(self) -> bytes:
        return self.proof.pi.state_in_hash

    def state_out_hash(self) -> bytes:
        return self.proof.pi.state_out_hash

    def to_bytes(self) -> bytes:
        """step.rs:79-152."""
        p, core = self.proof, self.pi_core
        out = bytearray(MAGIC)
        out += p.meta.lambda_bits.to_bytes(4, "little")
        out += p.header.suite_id + bytes(core.program_id) + bytes(core.program_commitment) + bytes(core.merkle_root)
        out += core.feature_mask.to_bytes(8, "little")
        out += encode_main_args(core.main_args)
        out += self.vm_usage_mask.to_bytes(4, "little") + self.ram_delta_clk_bits.to_bytes(4, "little")
        for lane in self.rom_acc:
            out += fe_to_bytes_fold(lane)
        out += p.pi.segment_index.to_bytes(4, "little") + p.pi.segments_total.to_bytes(4, "little") + p.pi.pc_init
        out += p.pi.boundary_bytes()
        out += len(p.inner_bytes).to_bytes(4, "little") + p.inner_bytes
        return bytes(out)

    @staticmethod
    def from_bytes(data: bytes) -> "StepProof":
        """step.rs:154-495, same checks and messages."""
        data = bytes(data)
        if len(data) < len(MAGIC):
            raise InvalidInput("step proof too short to contain magic header")
        if data[:len(MAGIC)] != MAGIC:
            raise InvalidInput("invalid step proof magic tag")
        cur = [len(MAGIC)]

        def take(n, msg):
            if len(data) < cur[0] + n:
                raise InvalidInput(msg)
            out = data[cur[0]:cur[0] + n]
            cur[0] += n
            return out

        def take_int(n, what):
            return int.from_bytes(take(n, "step proof truncated before " + what), "little")

        def take_32(what):
            known = what in BOUNDARY_FIELDS or what in ("suite_id", "program_id", "program_commitment", "merkle_root")
            return take(32, f"step proof truncated before {what} bytes" if known else "step proof truncated before 32-byte field")

        lambda_bits = take_int(4, "lambda_bits")
        suite_id, program_id = take_32("suite_id"), take_32("program_id")
        program_commitment, merkle_root = take_32("program_commitment"), take_32("merkle_root")
        feature_mask = take_int(8, "feature_mask")
        main_args = []
        for _ in range(take_int(4, "main_args length")):
            tag = take(1, "step proof truncated in VmArg tag")[0]
            if tag == 0:
                main_args.append(("u64", int.from_bytes(take(8, "step proof truncated in VmArg::U64"), "little")))
            elif tag == 1:
                main_args.append(("u128", int.from_bytes(take(16, "step proof truncated in VmArg::U128"), "little")))
            elif tag == 2:
                main_args.append(("bytes32", take(32, "step proof truncated in VmArg::Bytes32")))
            else:
                raise InvalidInput("invalid VmArg tag in step proof encoding")
        vm_usage_mask = take_int(4, "vm_usage_mask")
        ram_delta_clk_bits = take_int(4, "ram_delta_clk_bits")
        rom_acc = tuple(fe_from_bytes_fold(take_32(f"rom_acc_{i}")) for i in range(3))
        segment_index = take_int(4, "segment_index")
        segments_total = take_int(4, "segments_total")
        pc_init = take_32("pc_init")
        boundaries = {f: take_32(f) for f in BOUNDARY_FIELDS}
        inner_len = take_int(4, "inner proof length")
        inner_bytes = take(inner_len, "step proof truncated before inner proof bytes")
        inner = InnerProof.parse(inner_bytes)

        core = A.PublicInputs(program_id=program_id, program_commitment=program_commitment, merkle_root=merkle_root,
                              main_args=main_args, feature_mask=feature_mask)
        # pi_len is recomputed from the AIR public inputs with zeroed boundary values (step.rs:416-432)
        air_pi = A.AirPublicInputs(core=core, rom_acc=rom_acc, vm_usage_mask=vm_usage_mask, ram_delta_clk_bits=ram_delta_clk_bits)
        meta = StepMeta.from_env(inner.trace_len, inner.blowup, inner.num_queries, lambda_bits, len(air_pi.to_elements()))
        if segments_total <= 1:
            zl1 = Zl1Proof.new_single_segment(suite_id, meta, core, pc_init, boundaries, inner_bytes)
        else:
            zl1 = Zl1Proof.new_multi_segment(suite_id, meta, core, segment_index, segments_total, pc_init, boundaries, inner_bytes)
        return StepProof(zl1, core, rom_acc, vm_usage_mask, ram_delta_clk_bits)


# ---------------------------------------------------------------------------------------------------------------
# what the aggregation layer derives from a batch of step proofs
# ---------------------------------------------------------------------------------------------------------------
def children_root(suite_id: bytes, children: Sequence[Tuple[bytes, bytes]]) -> bytes:
    """agg/child.rs:853-895 over (step_digest, root_trace) pairs: sorted leaves, odd nodes paired with themselves."""
    if not children:
        return _Z32
    items = sorted(fe_to_bytes_fold(x) for x in hash_two_lanes_batch(suite_id, [(fold_bytes32_to_fe(d), fold_bytes32_to_fe(t))
                                                                                for d, t in children]))
    layer = [fold_bytes32_to_fe(b) for b in items]
    while len(layer) > 1:
        layer = hash_two_lanes_batch(suite_id, [(layer[i], layer[i + 1] if i + 1 < len(layer) else layer[i])
                                                for i in range(0, len(layer), 2)])
    return fe_to_bytes_fold(layer[0])


def children_root_from_steps(steps: Sequence[StepProof]) -> bytes:
    """lib.rs:417-427: the suite of the first child keys the tree."""
    if not steps:
        return _Z32
    return children_root(steps[0].proof.header.suite_id, [(s.digest(), s.proof.root_trace) for s in steps])


def verify_chain(steps: Sequence[StepProof]) -> None:
    """State chaining across consecutive segments (prove.rs:1040-1046 / vm/trace/mod.rs:353-359; skipped by the
    reference's parallel mode while proving, checked by the aggregation layer afterwards)."""
    for prev, cur in zip(steps, steps[1:]):
        if prev.state_out_hash() != cur.state_in_hash():
            raise InvalidInput("prev_state.state_out_hash must match segment state_in_hash")
        if cur.proof.pi.segment_index != prev.proof.pi.segment_index + 1:
            raise InvalidInput("segment indices must be consecutive")
