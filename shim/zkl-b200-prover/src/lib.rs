//! GPU replacements for the three hook types of zk-lisp's Winterfell prover
//! (`zk-lisp-proof-winterfell/src/prove.rs:425-517`):
//!
//! | reference (CPU default)                                   | here                       | C ABI                                         |
//! |-----------------------------------------------------------|----------------------------|-----------------------------------------------|
//! | `DefaultTraceLde` (`prove.rs:432-433`, built at :462)     | [`GpuTraceLde`]            | `zkl_trace_lde_commit`, `zkl_trace_query`     |
//! | `DefaultConstraintEvaluator` (`:434-435`, built at :483)  | [`GpuConstraintEvaluator`] | `zkl_eval_constraints`                        |
//! | `DefaultConstraintCommitment` (`:436-437`, built at :502) | [`GpuConstraintCommitment`]| `zkl_constraint_commit`, `zkl_constraint_query` |
//!
//! plus [`prove_segment_full`] for the stages that live in the provided method `Prover::prove` (OOD, DEEP, FRI, grinding,
//! queries), which the hooks cannot reach. `HashFn`, `VC = MerkleTree<HashFn>` and `RandomCoin` stay the reference's types: the
//! verifier is unchanged.
//!
//! Never compiled in the repository's build image (no Rust toolchain there); written against winterfell 0.13.1.
use std::ffi::CStr;
use std::marker::PhantomData;
use std::sync::Arc;

use winterfell::crypto::{BatchMerkleProof, ElementHasher, Hasher, MerkleTree};
use winterfell::math::fields::f128::BaseElement as BE;
use winterfell::math::{FieldElement, StarkField};
use winterfell::matrix::ColMatrix;
use winterfell::{
    Air, CompositionPoly, CompositionPolyTrace, ConstraintCommitment, ConstraintCompositionCoefficients, ConstraintEvaluator,
    EvaluationFrame, PartitionOptions, Proof, ProofOptions, StarkDomain, TraceInfo, TraceLde, TracePolyTable,
};
use zk_lisp_proof::pi::VmArg;
use zk_lisp_proof_winterfell::poseidon::hasher::{PoseidonDigest, PoseidonHasher};
use zk_lisp_proof_winterfell::vm::air::ZkLispAir;
use zk_lisp_proof_winterfell::{utils, AirPublicInputs};
use zkl_b200_sys as sys;

pub type H = PoseidonHasher<BE>;
pub type VC = MerkleTree<H>;

// ------------------------------------------------------------------------------------------------------------------
// context: one CUDA stream pinned to one device per prover thread (prove.rs:1020-1048 runs one prover per rayon worker)
// ------------------------------------------------------------------------------------------------------------------
pub struct GpuCtxInner(*mut sys::zkl_ctx);
unsafe impl Send for GpuCtxInner {}
unsafe impl Sync for GpuCtxInner {} // hook objects are only used from the proving thread; Sync is what `TraceLde: Sync` demands
impl Drop for GpuCtxInner {
    fn drop(&mut self) {
        unsafe { sys::zkl_ctx_destroy(self.0) }
    }
}
#[derive(Clone)]
pub struct GpuCtx(Arc<GpuCtxInner>);

impl GpuCtx {
    /// `device = segment_index % num_gpus` (the segment index is at hand at `prove.rs:1060`).
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sys::zkl_ctx_create(device, &mut h) };
        assert_eq!(rc, 0, "zkl_ctx_create: {}", last_error(std::ptr::null()));
        // commitment / coin suite, suite_id = [0;32] (poseidon/hasher.rs:23): the reference's own derivation
        let ps = zk_lisp_proof_winterfell::poseidon::get_poseidon_suite(&[0u8; 32]);
        let mut s: sys::ZklSuite = unsafe { std::mem::zeroed() };
        for i in 0..2 {
            s.dom[i] = fe(ps.dom[i]);
        }
        for i in 0..12 {
            for j in 0..12 {
                s.mds[i][j] = fe(ps.mds[i][j]);
            }
        }
        for (r, row) in ps.rc.iter().enumerate().take(sys::ZKL_ROUNDS) {
            for j in 0..12 {
                s.rc[r][j] = fe(row[j]);
            }
        }
        let ctx = GpuCtx(Arc::new(GpuCtxInner(h)));
        ctx.check(unsafe { sys::zkl_suite_load(h, &s) });
        ctx
    }
    pub fn raw(&self) -> *mut sys::zkl_ctx {
        self.0 .0
    }
    /// CUDA failures come back as a status + text; turned into a panic with that text so that `catch_unwind`
    /// (`prove.rs:225-227`) reports it like any other prover failure.
    pub fn check(&self, rc: i32) {
        if rc != 0 {
            panic!("zkl_b200: {}", last_error(self.raw()));
        }
    }
}
fn last_error(ctx: *const sys::zkl_ctx) -> String {
    unsafe { CStr::from_ptr(sys::zkl_last_error(ctx)).to_string_lossy().into_owned() }
}
#[inline]
fn fe(x: BE) -> sys::ZklFe {
    sys::ZklFe::from(x.as_int())
}
#[inline]
fn digest32(b: [u8; 32]) -> PoseidonDigest {
    PoseidonDigest::from(b)
}

pub fn zkl_options(o: &ProofOptions, p: PartitionOptions) -> sys::ZklOptions {
    // PartitionOptions has no public getters in 0.13; zk-lisp builds it from (num_partitions, hash_rate) at prove.rs:1106-1107
    // via utils::select_partitions_for_trace, and ProofOptions::partition_options() returns the same pair — carried here by the
    // caller (GpuHooks::partitions).
    let _ = p;
    sys::ZklOptions {
        num_queries: o.num_queries() as u32,
        blowup: o.blowup_factor() as u32,
        grinding: o.grinding_factor(),
        num_partitions: 1,
        hash_rate: 0,
        fri_folding: o.to_fri_options().folding_factor() as u32,
        fri_remainder_max_degree: o.to_fri_options().remainder_max_degree() as u32,
        field_extension: o.field_extension() as u32, // FieldExtension::None = 1, Quadratic = 2
    }
}

// ------------------------------------------------------------------------------------------------------------------
// public inputs: AirPublicInputs (lib.rs:74-96) -> zkl_public_inputs; per-program constants -> zkl_program
// ------------------------------------------------------------------------------------------------------------------
pub fn zkl_public_inputs(pi: &AirPublicInputs) -> sys::ZklPublicInputs {
    let mut out: sys::ZklPublicInputs = unsafe { std::mem::zeroed() };
    out.feature_mask = pi.core.feature_mask;
    out.segment_feature_mask = pi.segment_feature_mask;
    out.merkle_root = pi.core.merkle_root;
    out.vm_expected_bytes = pi.core.vm_expected_bytes;
    out.vm_out_reg = pi.core.vm_out_reg as u32;
    out.vm_out_row = pi.core.vm_out_row;
    let slots = utils::encode_main_args_to_slots(&pi.core.main_args);
    assert!(slots.len() <= sys::ZKL_MAX_MAIN_SLOTS, "too many main_args slots");
    out.num_main_slots = slots.len() as u32;
    for (j, v) in slots.iter().enumerate() {
        out.main_slots[j] = fe(*v);
    }
    out.vm_usage_mask = pi.vm_usage_mask;
    out.ram_delta_clk_bits = pi.ram_delta_clk_bits;
    out.pc_init = fe(pi.pc_init);
    out.ram_gp_unsorted_in = fe(pi.ram_gp_unsorted_in);
    out.ram_gp_unsorted_out = fe(pi.ram_gp_unsorted_out);
    out.ram_gp_sorted_in = fe(pi.ram_gp_sorted_in);
    out.ram_gp_sorted_out = fe(pi.ram_gp_sorted_out);
    for i in 0..3 {
        out.rom_s_in[i] = fe(pi.rom_s_in[i]);
        out.rom_s_out[i] = fe(pi.rom_s_out[i]);
    }
    let _: Option<&VmArg> = pi.core.main_args.first();
    out
}
/// Per-program constants, derived by the library exactly as poseidon/mod.rs / utils.rs / commit.rs do (tests/kat.rs checks
/// them against the reference's own functions).
pub fn zkl_program(program_id: &[u8; 32], program_commitment: &[u8; 32]) -> Box<sys::ZklProgram> {
    let mut p: Box<sys::ZklProgram> = Box::new(unsafe { std::mem::zeroed() });
    let rc = unsafe { sys::zkl_program_derive(program_id.as_ptr(), program_commitment.as_ptr(), &mut *p) };
    assert_eq!(rc, 0, "zkl_program_derive failed");
    p
}

fn column_major(m: &ColMatrix<BE>) -> Vec<BE> {
    let (w, n) = (m.num_cols(), m.num_rows());
    let mut v = Vec::with_capacity(w * n);
    for c in 0..w {
        v.extend_from_slice(m.get_column(c));
    }
    v
}
fn batch_proof(depth: u8, list_len: &[u32], nodes: &[u8]) -> BatchMerkleProof<H> {
    let mut lists = Vec::with_capacity(list_len.len());
    let mut off = 0usize;
    for &l in list_len {
        let mut v = Vec::with_capacity(l as usize);
        for _ in 0..l {
            let mut d = [0u8; 32];
            d.copy_from_slice(&nodes[off..off + 32]);
            v.push(digest32(d));
            off += 32;
        }
        lists.push(v);
    }
    BatchMerkleProof { nodes: lists, depth }
}
/// rows (k x width, row-major) + the batch opening of a committed matrix
fn query_rows<E: FieldElement<BaseField = BE>>(
    ctx: &GpuCtx, width: usize, positions: &[usize],
    f: impl FnOnce(*const u32, usize, *mut sys::ZklFe, *mut sys::ZklBatchProof) -> i32,
) -> winterfell::Queries {
    let k = positions.len();
    let pos: Vec<u32> = positions.iter().map(|&p| p as u32).collect();
    let mut rows = vec![BE::ZERO; k * width];
    let mut list_len = vec![0u32; k.max(1)];
    let mut nodes = vec![0u8; k.max(1) * 64 * 32];
    let mut bp = sys::ZklBatchProof { depth: 0, num_lists: 0, list_len: list_len.as_mut_ptr(), nodes: nodes.as_mut_ptr(), total_nodes: 0 };
    ctx.check(f(pos.as_ptr(), k, rows.as_mut_ptr().cast(), &mut bp));
    let proof = batch_proof(bp.depth, &list_len[..bp.num_lists as usize], &nodes[..bp.total_nodes as usize * 32]);
    let values: Vec<Vec<E>> = rows.chunks(width).map(|r| r.iter().map(|&x| E::from(x)).collect()).collect();
    winterfell::Queries::new::<H, E, VC>(proof, values)
}

// ------------------------------------------------------------------------------------------------------------------
// hook 1: TraceLde (prove.rs:454-474)
// ------------------------------------------------------------------------------------------------------------------
pub struct GpuTraceLde<E: FieldElement<BaseField = BE>> {
    ctx: GpuCtx,
    handle: *mut sys::zkl_trace_lde,
    root: PoseidonDigest,
    info: TraceInfo,
    blowup: usize,
    _e: PhantomData<E>,
}
unsafe impl<E: FieldElement<BaseField = BE>> Send for GpuTraceLde<E> {}
unsafe impl<E: FieldElement<BaseField = BE>> Sync for GpuTraceLde<E> {}
impl<E: FieldElement<BaseField = BE>> Drop for GpuTraceLde<E> {
    fn drop(&mut self) {
        unsafe { sys::zkl_trace_lde_free(self.ctx.raw(), self.handle) }
    }
}
impl<E: FieldElement<BaseField = BE>> GpuTraceLde<E> {
    /// What `DefaultTraceLde::new(trace_info, main_trace, domain, partition_option)` does on the CPU.
    pub fn new(ctx: &GpuCtx, trace_info: &TraceInfo, main_trace: &ColMatrix<BE>, domain: &StarkDomain<BE>, opts: &sys::ZklOptions) -> (Self, TracePolyTable<E>) {
        let (w, n) = (main_trace.num_cols(), main_trace.num_rows());
        let cols = column_major(main_trace);
        let mut root = [0u8; 32];
        let mut polys = vec![BE::ZERO; w * n];
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { sys::zkl_trace_lde_commit(ctx.raw(), cols.as_ptr().cast(), w, n, opts, root.as_mut_ptr(), polys.as_mut_ptr().cast(), &mut h) });
        let table = TracePolyTable::new(ColMatrix::new(polys.chunks(n).map(|c| c.to_vec()).collect()));
        (GpuTraceLde { ctx: ctx.clone(), handle: h, root: digest32(root), info: trace_info.clone(), blowup: domain.trace_to_lde_blowup(), _e: PhantomData }, table)
    }
    pub fn handle(&self) -> *const sys::zkl_trace_lde {
        self.handle
    }
}
impl<E: FieldElement<BaseField = BE>> TraceLde<E> for GpuTraceLde<E> {
    type HashFn = H;
    type VC = VC;
    fn get_main_trace_commitment(&self) -> PoseidonDigest {
        self.root
    }
    fn set_aux_trace(&mut self, _aux: &ColMatrix<E>, _domain: &StarkDomain<BE>) -> (ColMatrix<E>, PoseidonDigest) {
        unreachable!("zk-lisp has no auxiliary trace segment (preflight.rs:89)")
    }
    fn read_main_trace_frame_into(&self, lde_step: usize, frame: &mut EvaluationFrame<BE>) {
        // only the stock CPU evaluator walks frames; the GPU evaluator reads the device-resident LDE directly
        let w = self.info.main_trace_width();
        let (mut cur, mut next) = (vec![BE::ZERO; w], vec![BE::ZERO; w]);
        self.ctx.check(unsafe { sys::zkl_trace_lde_read_frame(self.ctx.raw(), self.handle, lde_step, cur.as_mut_ptr().cast(), next.as_mut_ptr().cast()) });
        frame.current_mut().copy_from_slice(&cur);
        frame.next_mut().copy_from_slice(&next);
    }
    fn read_aux_trace_frame_into(&self, _lde_step: usize, _frame: &mut EvaluationFrame<E>) {
        unreachable!("no auxiliary trace segment")
    }
    fn query(&self, positions: &[usize]) -> Vec<winterfell::Queries> {
        let w = self.info.main_trace_width();
        vec![query_rows::<BE>(&self.ctx, w, positions, |p, k, rows, bp| unsafe { sys::zkl_trace_query(self.ctx.raw(), self.handle, p, k, rows, bp) })]
    }
    fn trace_len(&self) -> usize {
        self.info.length()
    }
    fn blowup(&self) -> usize {
        self.blowup
    }
    fn trace_info(&self) -> &TraceInfo {
        &self.info
    }
}

// ------------------------------------------------------------------------------------------------------------------
// hook 2: ConstraintEvaluator (prove.rs:476-492)
// ------------------------------------------------------------------------------------------------------------------
pub struct GpuConstraintEvaluator<'a, E: FieldElement<BaseField = BE>> {
    ctx: GpuCtx,
    air: &'a ZkLispAir,
    program: Box<sys::ZklProgram>,
    pi: sys::ZklPublicInputs,
    opts: sys::ZklOptions,
    coeffs: ConstraintCompositionCoefficients<E>,
    /// device handle of the trace LDE the prover built in hook 1 (the trait's `evaluate` only sees `&T: TraceLde`)
    lde: *const sys::zkl_trace_lde,
}
impl<'a, E: FieldElement<BaseField = BE>> GpuConstraintEvaluator<'a, E> {
    pub fn new(ctx: &GpuCtx, air: &'a ZkLispAir, air_pi: &AirPublicInputs, opts: sys::ZklOptions, coeffs: ConstraintCompositionCoefficients<E>, lde: *const sys::zkl_trace_lde) -> Self {
        GpuConstraintEvaluator {
            ctx: ctx.clone(), air, program: zkl_program(&air_pi.core.program_id, &air_pi.core.program_commitment),
            pi: zkl_public_inputs(air_pi), opts, coeffs, lde,
        }
    }
}
impl<'a, E: FieldElement<BaseField = BE>> ConstraintEvaluator<'a, E> for GpuConstraintEvaluator<'a, E> {
    type Air = ZkLispAir;
    fn evaluate<T: TraceLde<E>>(self, trace: &T, domain: &'a StarkDomain<BE>) -> CompositionPolyTrace<E> {
        assert_eq!(E::EXTENSION_DEGREE, 1, "step proofs use FieldExtension::None (prove.rs:967)");
        let (w, n) = (trace.trace_info().main_trace_width(), trace.trace_len());
        // the AIR instance as the library derives it from the public inputs (zkl_air_build) ...
        let cap = (n / 32) * 141 + 64;
        let mut assertions = vec![sys::ZklAssertion { column: 0, step: 0, value: sys::ZklFe::default() }; cap];
        let mut seed = vec![sys::ZklFe::default(); 64];
        let (mut na, mut ns) = (cap, 64usize);
        let mut params: sys::ZklAirParams = unsafe { std::mem::zeroed() };
        let mut err = [0i8; 256];
        let rc = unsafe { sys::zkl_air_build(&*self.program, &self.pi, w, n, &self.opts, &mut params, assertions.as_mut_ptr(), &mut na, seed.as_mut_ptr(), &mut ns, err.as_mut_ptr().cast(), 256) };
        assert_eq!(rc, 0, "zkl_air_build: {}", unsafe { CStr::from_ptr(err.as_ptr().cast()) }.to_string_lossy());
        // ... must be the instance Winterfell holds: same constraint counts, same CE domain
        let ctx_air = self.air.context();
        assert_eq!(params.num_transition as usize, ctx_air.num_main_transition_constraints());
        assert_eq!(params.num_assertions as usize, ctx_air.num_assertions());
        assert_eq!(params.ce_blowup as usize * n, domain.ce_domain_size());
        // transition coefficients in evaluate_transition order, then boundary ones in Winterfell's sorted-assertion order
        let coeffs: Vec<BE> = self.coeffs.transition.iter().chain(self.coeffs.boundary.iter()).map(|c| c.base_element(0)).collect();
        assert_eq!(coeffs.len(), (params.num_transition + params.num_assertions) as usize);
        let mut out = vec![BE::ZERO; domain.ce_domain_size()];
        self.ctx.check(unsafe { sys::zkl_eval_constraints(self.ctx.raw(), self.lde, &params, coeffs.as_ptr().cast(), out.as_mut_ptr().cast()) });
        CompositionPolyTrace::new(out.into_iter().map(E::from).collect())
    }
}

// ------------------------------------------------------------------------------------------------------------------
// hook 3: ConstraintCommitment (prove.rs:494-516)
// ------------------------------------------------------------------------------------------------------------------
pub struct GpuConstraintCommitment<E: FieldElement<BaseField = BE>> {
    ctx: GpuCtx,
    handle: *mut sys::zkl_constraint_commitment,
    root: PoseidonDigest,
    num_cols: usize,
    _e: PhantomData<E>,
}
impl<E: FieldElement<BaseField = BE>> Drop for GpuConstraintCommitment<E> {
    fn drop(&mut self) {
        unsafe { sys::zkl_constraint_commitment_free(self.ctx.raw(), self.handle) }
    }
}
impl<E: FieldElement<BaseField = BE>> GpuConstraintCommitment<E> {
    /// What `DefaultConstraintCommitment::new(composition_poly_trace, num_cols, domain, partition_options)` does on the CPU.
    /// Winterfell 0.13 offers no public constructor of `CompositionPoly` from column coefficients, so the (cheap, 8n-point)
    /// interpolation that yields the `CompositionPoly` stays on `CompositionPoly::new`; the GPU does the same interpolation for
    /// its own LDE + row hashing + Merkle tree, which is where the time goes (7 x 65 536 elements, 131 k permutations).
    pub fn new(ctx: &GpuCtx, trace: CompositionPolyTrace<E>, num_cols: usize, domain: &StarkDomain<BE>, opts: &sys::ZklOptions) -> (Self, CompositionPoly<E>) {
        assert_eq!(E::EXTENSION_DEGREE, 1);
        let n = domain.trace_length();
        let evals: Vec<BE> = trace.clone().into_inner().iter().map(|e| e.base_element(0)).collect();
        let mut root = [0u8; 32];
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { sys::zkl_constraint_commit(ctx.raw(), evals.as_ptr().cast(), n, evals.len() / n, num_cols, opts, root.as_mut_ptr(), std::ptr::null_mut(), &mut h) });
        let poly = CompositionPoly::new(trace, domain, num_cols);
        (GpuConstraintCommitment { ctx: ctx.clone(), handle: h, root: digest32(root), num_cols, _e: PhantomData }, poly)
    }
}
impl<E: FieldElement<BaseField = BE>> ConstraintCommitment<E> for GpuConstraintCommitment<E> {
    type HashFn = H;
    type VC = VC;
    fn commitment(&self) -> PoseidonDigest {
        self.root
    }
    fn query(self, positions: &[usize]) -> winterfell::Queries {
        query_rows::<E>(&self.ctx, self.num_cols, positions, |p, k, rows, bp| unsafe { sys::zkl_constraint_query(self.ctx.raw(), self.handle, p, k, rows, bp) })
    }
}

// ------------------------------------------------------------------------------------------------------------------
// whole proof (mode `full`): replaces the call `prover.prove(trace)` at prove.rs:225
// ------------------------------------------------------------------------------------------------------------------
/// `trace`: the segment's TraceTable columns; `air_pi` = `Prover::get_pub_inputs(trace)` (prove.rs:441-450).
/// Returns the `winterfell::Proof` parsed from the bytes the device wrote — the stock verifier takes it from there.
pub fn prove_segment_full(ctx: &GpuCtx, main_trace: &ColMatrix<BE>, air_pi: &AirPublicInputs, options: &ProofOptions, partitions: (usize, usize)) -> Result<Proof, String> {
    let (w, n) = (main_trace.num_cols(), main_trace.num_rows());
    let cols = column_major(main_trace);
    let program = zkl_program(&air_pi.core.program_id, &air_pi.core.program_commitment);
    let pi = zkl_public_inputs(air_pi);
    let mut opts = zkl_options(options, PartitionOptions::new(partitions.0, partitions.1));
    opts.num_partitions = partitions.0 as u32;
    opts.hash_rate = partitions.1 as u32;
    let mut buf = vec![0u8; 1 << 20];
    loop {
        let mut len = buf.len();
        let rc = unsafe { sys::zkl_prove_segment_pi(ctx.raw(), cols.as_ptr().cast(), w, n, &*program, &pi, &opts, buf.as_mut_ptr(), &mut len, std::ptr::null_mut()) };
        match rc {
            0 => return Proof::from_bytes(&buf[..len]).map_err(|e| e.to_string()),
            sys::ZKL_ERR_BUFFER_TOO_SMALL => buf.resize(len + len / 4, 0),
            _ => return Err(last_error(ctx.raw())), // incl. ZKL_ERR_CONSTRAINTS_VIOLATED: Winterfell's debug builds panic on the same condition
        }
    }
}

/// What the patched `ZkWinterfellProver` keeps per proof (see zk-lisp-proof-winterfell.patch).
pub struct GpuHooks {
    pub ctx: GpuCtx,
    pub partitions: (usize, usize),
    pub air_pi: std::sync::OnceLock<AirPublicInputs>,
    pub lde: std::sync::atomic::AtomicPtr<sys::zkl_trace_lde>,
}
impl GpuHooks {
    pub fn new(device: i32, partitions: (usize, usize)) -> Self {
        GpuHooks { ctx: GpuCtx::new(device), partitions, air_pi: std::sync::OnceLock::new(), lde: std::sync::atomic::AtomicPtr::new(std::ptr::null_mut()) }
    }
    pub fn options(&self, o: &ProofOptions) -> sys::ZklOptions {
        let mut z = zkl_options(o, PartitionOptions::new(self.partitions.0, self.partitions.1));
        z.num_partitions = self.partitions.0 as u32;
        z.hash_rate = self.partitions.1 as u32;
        z
    }
}

#[allow(unused)]
fn _assert_traits() {
    fn is_hasher<T: ElementHasher<BaseField = BE> + Hasher<Digest = PoseidonDigest>>() {}
    is_hasher::<H>();
    let _ = BE::MODULUS;
}

// ------------------------------------------------------------------------------------------------------------------
// the aggregation proof (prove.rs:526-581 hooks, :629-702 prove_agg_proof): same entry point, air_kind = ZKL_AIR_AGG
// ------------------------------------------------------------------------------------------------------------------
/// `trace`: the 31-column aggregation trace (`build_agg_trace_from_transcripts`, agg/trace.rs:155-247); `agg_pi` the aggregation
/// public inputs; `options` as `prove_agg_proof` builds them (FieldExtension::Quadratic from 128 bits up). The coin seed is
/// Context::to_elements() ++ AggAirPublicInputs::to_elements() (agg/pi.rs:176-217).
pub fn prove_agg_full(ctx: &GpuCtx, trace: &ColMatrix<BE>, agg_pi: &zk_lisp_proof_winterfell::agg::pi::AggAirPublicInputs, options: &ProofOptions,
                      partitions: (usize, usize), num_constraints: usize, trace_info: &TraceInfo) -> Result<Proof, String> {
    use winterfell::math::ToElements;
    let (w, n) = (trace.num_cols(), trace.num_rows());
    let cols = column_major(trace);
    let mut params: sys::ZklAirParams = unsafe { std::mem::zeroed() };
    let mut assertions = [sys::ZklAssertion { column: 0, step: 0, value: sys::ZklFe::default() }; 5];
    let rc = unsafe { sys::zkl_agg_air_params(n, agg_pi.v_units_total, agg_pi.children_count, &mut params, assertions.as_mut_ptr()) };
    if rc != 0 {
        return Err("zkl_agg_air_params: trace length must be a power of two >= 8".into());
    }
    assert_eq!((params.num_transition + params.num_assertions) as usize, num_constraints);
    let mut seed: Vec<BE> = winterfell::Context::new::<BE>(trace_info.clone(), options.clone(), num_constraints).to_elements();
    seed.extend(agg_pi.to_elements());
    let mut opts = zkl_options(options, PartitionOptions::new(partitions.0, partitions.1));
    opts.num_partitions = partitions.0 as u32;
    opts.hash_rate = partitions.1 as u32;
    let mut buf = vec![0u8; 1 << 20];
    let mut len = buf.len();
    let rc = unsafe { sys::zkl_prove_segment(ctx.raw(), cols.as_ptr().cast(), w, n, &params, seed.as_ptr().cast(), seed.len(), &opts, buf.as_mut_ptr(), &mut len, std::ptr::null_mut()) };
    if rc != 0 {
        return Err(last_error(ctx.raw()));
    }
    Proof::from_bytes(&buf[..len]).map_err(|e| e.to_string())
}

// ------------------------------------------------------------------------------------------------------------------
// one large segment over the GPUs of a box (include/zkl_b200.h: zkl_comm)
// ------------------------------------------------------------------------------------------------------------------
/// This rank's share of the column-sharded proof of ONE segment: `trace` is the whole `ColMatrix` on the host (every rank of a
/// single-process prover sees it; a multi-process prover passes a matrix that holds valid data in its own columns only), the rank
/// uploads the columns `zkl_shard_columns` assigns it and proves with the caller's collectives. Every rank returns the same
/// `Proof`, byte-identical to `prove_segment_full` on one GPU. `partitions` must give at least `comm.world` partitions.
pub fn prove_segment_sharded(ctx: &GpuCtx, comm: &sys::ZklComm, trace: &ColMatrix<BE>, program: &sys::ZklProgram, pi: &sys::ZklPublicInputs,
                             options: &ProofOptions, partitions: (usize, usize)) -> Result<Proof, String> {
    let (w, n) = (trace.num_cols(), trace.num_rows());
    let mut opts = zkl_options(options, PartitionOptions::new(partitions.0, partitions.1));
    opts.num_partitions = partitions.0 as u32;
    opts.hash_rate = partitions.1 as u32;
    let (mut first, mut count) = (0usize, 0usize);
    if unsafe { sys::zkl_shard_columns(&opts, w, comm.rank, comm.world, &mut first, &mut count) } != 0 {
        return Err("column sharding needs PartitionOptions with at least one whole partition per rank".into());
    }
    let mut local: Vec<BE> = Vec::with_capacity(count * n);
    for c in first..first + count {
        local.extend_from_slice(trace.get_column(c));
    }
    let bytes = local.len() * std::mem::size_of::<BE>();
    let mut dev: *mut std::ffi::c_void = std::ptr::null_mut();
    if unsafe { sys::zkl_dev_alloc(ctx.raw(), bytes, &mut dev) } != 0 || unsafe { sys::zkl_dev_upload(ctx.raw(), dev, local.as_ptr().cast(), bytes) } != 0 {
        return Err(last_error(ctx.raw()));
    }
    let mut buf = vec![0u8; 4 << 20];
    let mut len = buf.len();
    let rc = unsafe {
        sys::zkl_prove_segment_sharded_pi_dev(ctx.raw(), dev, first, count, w, n, program, pi, &opts, comm, buf.as_mut_ptr(), &mut len, std::ptr::null_mut())
    };
    unsafe { sys::zkl_dev_free(ctx.raw(), dev) };
    if rc != 0 {
        return Err(last_error(ctx.raw()));
    }
    Proof::from_bytes(&buf[..len]).map_err(|e| e.to_string())
}
