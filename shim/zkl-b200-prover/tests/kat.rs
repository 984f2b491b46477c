//! Pins this repository's byte-level assumptions against the REAL crates (winterfell 0.13.1, zk-lisp v0.16.0).
//! Every `[WF-mem]` item of DESIGN.md §3 has its own test; a failure names the assumption that is wrong.
//! Fixtures: tests/golden/{zkl_kat.json, segment64_trace.bin, segment64_pi.json, segment64_proof.bin} of the zk-lisp_b200 checkout.
//! Tests whose name starts with `gpu_` need libzkl_b200.so and a B200; the others need only the reference crates.
use serde_json::Value;
use winterfell::crypto::{DefaultRandomCoin, ElementHasher, Hasher, MerkleTree, RandomCoin};
use winterfell::math::fields::f128::BaseElement as BE;
use winterfell::math::{fft, FieldElement, StarkField, ToElements};
use winterfell::{AcceptableOptions, BatchingMethod, FieldExtension, Proof, ProofOptions, Trace, TraceTable};
use zk_lisp_proof::pi::PublicInputs;
use zk_lisp_proof_winterfell::poseidon::{self, hasher::PoseidonHasher};
use zk_lisp_proof_winterfell::vm::air::ZkLispAir;
use zk_lisp_proof_winterfell::{prove, utils, AirPublicInputs};

type H = PoseidonHasher<BE>;

fn root() -> String {
    std::env::var("ZKL_B200_ROOT").expect("ZKL_B200_ROOT")
}
fn golden(name: &str) -> Vec<u8> {
    std::fs::read(format!("{}/tests/golden/{name}", root())).unwrap()
}
fn kat() -> Value {
    serde_json::from_slice(&golden("zkl_kat.json")).unwrap()
}
fn fe_hex(x: BE) -> String {
    hex::encode(x.as_int().to_le_bytes())
}
fn fe_from_hex(s: &str) -> BE {
    let mut b = [0u8; 16];
    b.copy_from_slice(&hex::decode(s).unwrap());
    BE::new(u128::from_le_bytes(b))
}
/// x_i = LE128(blake3("zkl-b200/kat" || tag u32 LE || i u32 LE)[0..16]) mod p  (tests/golden/make_kat.py: gen_elements)
fn gen(tag: u32, n: usize) -> Vec<BE> {
    (0..n as u32)
        .map(|i| {
            let mut h = blake3::Hasher::new();
            h.update(b"zkl-b200/kat");
            h.update(&tag.to_le_bytes());
            h.update(&i.to_le_bytes());
            let mut b = [0u8; 16];
            b.copy_from_slice(&h.finalize().as_bytes()[..16]);
            BE::new(u128::from_le_bytes(b) % BE::MODULUS)
        })
        .collect()
}
fn digest_of(x: BE) -> <H as Hasher>::Digest {
    let mut b = [0u8; 32];
    b[..16].copy_from_slice(&x.as_int().to_le_bytes());
    b.into()
}
fn dhex(d: &<H as Hasher>::Digest) -> String {
    hex::encode(d.as_bytes())
}

// ---- suite constants (poseidon/mod.rs:52-217) ----
#[test]
fn suite_constants() {
    let k = kat();
    for name in ["hasher", "program"] {
        let s = &k["suites"][name];
        let mut sid = [0u8; 32];
        sid.copy_from_slice(&hex::decode(s["suite_id"].as_str().unwrap()).unwrap());
        let ps = poseidon::get_poseidon_suite(&sid);
        assert_eq!(fe_hex(ps.dom[0]), s["dom"][0].as_str().unwrap());
        assert_eq!(fe_hex(ps.dom[1]), s["dom"][1].as_str().unwrap());
        assert_eq!(fe_hex(ps.mds[0][0]), s["mds_0_0"].as_str().unwrap());
        assert_eq!(fe_hex(ps.mds[11][11]), s["mds_11_11"].as_str().unwrap());
        assert_eq!(fe_hex(ps.rc[0][0]), s["rc_0_0"].as_str().unwrap());
        assert_eq!(fe_hex(ps.rc[26][11]), s["rc_26_11"].as_str().unwrap());
        let mut flat = Vec::new();
        for x in ps.dom.iter().chain(ps.mds.iter().flatten()).chain(ps.rc.iter().take(27).flatten()) {
            flat.extend_from_slice(&x.as_int().to_le_bytes());
        }
        assert_eq!(hex::encode(blake3::hash(&flat).as_bytes()), s["blake3_of_dom_mds_rc"].as_str().unwrap());
    }
}

// ---- hasher entry points (poseidon/hasher.rs:57-140) ----
#[test]
fn hash_elements() {
    for c in kat()["hash_elements"].as_array().unwrap() {
        let e = gen(c["tag"].as_u64().unwrap() as u32, c["len"].as_u64().unwrap() as usize);
        assert_eq!(dhex(&H::hash_elements(&e)), c["digest"].as_str().unwrap(), "len {}", c["len"]);
    }
}
#[test]
fn merge_and_merge_with_int() {
    let k = kat();
    let ab = gen(3, 2);
    let (a, b) = (digest_of(ab[0]), digest_of(ab[1]));
    assert_eq!(dhex(&H::merge(&[a, b])), k["merge"]["digest"].as_str().unwrap());
    for c in k["merge_with_int"].as_array().unwrap() {
        assert_eq!(dhex(&H::merge_with_int(a, c["value"].as_u64().unwrap())), c["digest"].as_str().unwrap());
    }
}
#[test]
fn merge_many() {
    for c in kat()["merge_many"].as_array().unwrap() {
        let ds: Vec<_> = gen(4, c["count"].as_u64().unwrap() as usize).into_iter().map(digest_of).collect();
        assert_eq!(dhex(&H::merge_many(&ds)), c["digest"].as_str().unwrap());
    }
}
#[test]
fn hash_bytes() {
    for c in kat()["hash_bytes"].as_array().unwrap() {
        let data: Vec<u8> = (0..c["len"].as_u64().unwrap() as usize).map(|i| (i % 251) as u8).collect();
        assert_eq!(dhex(&H::hash(&data)), c["digest"].as_str().unwrap());
    }
}
#[test]
fn permutation_via_two_lanes() {
    // poseidon_hash_two_lanes (poseidon/mod.rs:255-291) = lane 0 of the permutation of [l, r, 0 x 8, dom0, dom1]
    let st = gen(1, 12);
    let _ = st; // the raw permutation is not public in the reference; hash_elements(len 10/20) above walks it 1-2 times
}

// ---- Merkle tree node order (winter-crypto MerkleTree::new) ----
#[test]
fn merkle_root_node_order() {
    let leaves: Vec<_> = gen(5, 8).into_iter().map(digest_of).collect();
    let tree = MerkleTree::<H>::new(leaves).unwrap();
    assert_eq!(dhex(tree.root()), kat()["merkle_root_8_leaves"]["root"].as_str().unwrap());
}

// ---- FFT output order, coset evaluation (winter-math fft) ----
#[test]
fn interpolation_and_coset_lde() {
    let k = kat();
    let mut col = gen(6, 8);
    let inv_tw = fft::get_inv_twiddles::<BE>(8);
    fft::interpolate_poly(&mut col, &inv_tw);
    let want: Vec<&str> = k["transform_8"]["poly"].as_array().unwrap().iter().map(|v| v.as_str().unwrap()).collect();
    assert_eq!(col.iter().map(|&x| fe_hex(x)).collect::<Vec<_>>(), want);
    let tw = fft::get_twiddles::<BE>(8);
    let lde = fft::evaluate_poly_with_offset(&col, &tw, BE::GENERATOR, 2);
    let want: Vec<&str> = k["transform_8"]["lde_blowup2_offset3"].as_array().unwrap().iter().map(|v| v.as_str().unwrap()).collect();
    assert_eq!(lde.iter().map(|&x| fe_hex(x)).collect::<Vec<_>>(), want);
}

// ---- the 64-row segment ----
struct Seg {
    trace: TraceTable<BE>,
    pi: PublicInputs,
    air_pi: AirPublicInputs,
    opts: ProofOptions,
    proof_bytes: Vec<u8>,
    j: Value,
}
fn bytes32(s: &str) -> [u8; 32] {
    let mut b = [0u8; 32];
    b.copy_from_slice(&hex::decode(s).unwrap());
    b
}
fn segment() -> Seg {
    let j: Value = serde_json::from_slice(&golden("segment64_pi.json")).unwrap();
    let (w, n) = (j["width"].as_u64().unwrap() as usize, j["n"].as_u64().unwrap() as usize);
    let raw = golden("segment64_trace.bin");
    assert_eq!(raw.len(), w * n * 16);
    let cols: Vec<Vec<BE>> = (0..w)
        .map(|c| (0..n).map(|r| { let o = (c * n + r) * 16; let mut b = [0u8; 16]; b.copy_from_slice(&raw[o..o + 16]); BE::new(u128::from_le_bytes(b)) }).collect())
        .collect();
    let trace = TraceTable::init(cols);
    let pi = PublicInputs {
        program_id: bytes32(j["program_id"].as_str().unwrap()),
        program_commitment: bytes32(j["program_commitment"].as_str().unwrap()),
        merkle_root: bytes32(j["merkle_root"].as_str().unwrap()),
        feature_mask: j["feature_mask"].as_u64().unwrap(),
        ..Default::default()
    };
    let f = |k: &str, i: usize| fe_from_hex(j[k][i].as_str().unwrap());
    let air_pi = AirPublicInputs {
        core: pi.clone(),
        segment_feature_mask: j["segment_feature_mask"].as_u64().unwrap(),
        rom_acc: [BE::ZERO; 3],
        pc_init: fe_from_hex(j["pc_init"].as_str().unwrap()),
        ram_gp_unsorted_in: f("ram_gp", 0), ram_gp_unsorted_out: f("ram_gp", 1), ram_gp_sorted_in: f("ram_gp", 2), ram_gp_sorted_out: f("ram_gp", 3),
        rom_s_in: [f("rom_s_in", 0), f("rom_s_in", 1), f("rom_s_in", 2)],
        rom_s_out: [f("rom_s_out", 0), f("rom_s_out", 1), f("rom_s_out", 2)],
        vm_usage_mask: j["vm_usage_mask"].as_u64().unwrap() as u32,
        ram_delta_clk_bits: j["ram_delta_clk_bits"].as_u64().unwrap() as u32,
    };
    let (np, hr) = (j["num_partitions"].as_u64().unwrap() as usize, j["hash_rate"].as_u64().unwrap() as usize);
    let opts = ProofOptions::new(j["num_queries"].as_u64().unwrap() as usize, j["blowup"].as_u64().unwrap() as usize, j["grinding"].as_u64().unwrap() as u32,
                                 FieldExtension::None, 2, 1, BatchingMethod::Linear, BatchingMethod::Linear).with_partitions(np, hr);
    Seg { trace, pi, air_pi, opts, proof_bytes: golden("segment64_proof.bin"), j }
}

/// [WF-mem] Context::to_elements + the order context ++ public inputs of the coin seed (agg/fs.rs:69-73)
#[test]
fn coin_seed_elements() {
    let s = segment();
    let air = <ZkLispAir as winterfell::Air>::new(s.trace.info().clone(), s.air_pi.clone(), s.opts.clone());
    let mut seed = winterfell::Air::context(&air).to_elements();
    seed.extend(s.air_pi.to_elements());
    let want: Vec<&str> = s.j["seed_elements"].as_array().unwrap().iter().map(|v| v.as_str().unwrap()).collect();
    assert_eq!(seed.iter().map(|&x| fe_hex(x)).collect::<Vec<_>>(), want, "Context::to_elements / AirPublicInputs::to_elements");
    assert_eq!(winterfell::Air::context(&air).num_constraints(), kat()["segment_proof"]["num_constraints"].as_u64().unwrap() as usize);
}
/// [WF-mem] Winterfell's assertion order: by (stride, first step, column)
#[test]
fn assertion_order() {
    let s = segment();
    let air = <ZkLispAir as winterfell::Air>::new(s.trace.info().clone(), s.air_pi.clone(), s.opts.clone());
    let mut a = winterfell::Air::get_assertions(&air);
    a.sort(); // the order BoundaryConstraints::new consumes them in (Assertion: Ord)
    for (i, w) in s.j["assertions_sorted_first_8"].as_array().unwrap().iter().enumerate() {
        assert_eq!((a[i].column() as u64, a[i].first_step() as u64), (w[0].as_u64().unwrap(), w[1].as_u64().unwrap()));
        assert_eq!(fe_hex(a[i].values()[0]), w[2].as_str().unwrap());
    }
}
/// [WF-mem] Proof / Queries / OodFrame / FriProof / BatchMerkleProof byte layouts, coefficient draw order, DEEP, FRI, grinding:
/// the stock verifier accepts the bytes this repository wrote.
#[test]
fn stock_verifier_accepts_committed_proof() {
    let s = segment();
    let proof = Proof::from_bytes(&s.proof_bytes).expect("Proof::from_bytes on zkl_prove_segment output");
    assert_eq!(proof.to_bytes(), s.proof_bytes, "Proof::to_bytes round trip");
    let acceptable = AcceptableOptions::OptionSet(vec![s.opts.clone()]);
    winterfell::verify::<ZkLispAir, H, DefaultRandomCoin<H>, MerkleTree<H>>(proof, s.air_pi.clone(), &acceptable).expect("winterfell::verify");
}
/// The definitive pin: the reference's CPU prover on the same trace yields the SAME bytes (proving is deterministic: no seed,
/// smallest nonce; zk-lisp-cli/src/prove.rs:27-31).
#[test]
fn stock_cpu_prover_yields_the_committed_bytes() {
    let s = segment();
    let prover = prove::ZkProver::new(s.opts.clone(), s.pi.clone(), [BE::ZERO; 3]);
    let proof = prover.prove(s.trace.clone()).expect("reference prover");
    let k = kat();
    assert_eq!(proof.to_bytes().len() as u64, k["segment_proof"]["proof_len"].as_u64().unwrap());
    assert_eq!(proof.to_bytes(), s.proof_bytes);
    assert_eq!(proof.pow_nonce, k["segment_proof"]["pow_nonce"].as_u64().unwrap());
}
/// [WF-mem] DefaultRandomCoin: seed = hash_elements, reseed = merge, draw = merge_with_int(seed, ++counter)
#[test]
fn random_coin_transcript() {
    let s = segment();
    let k = kat();
    let seed: Vec<BE> = s.j["seed_elements"].as_array().unwrap().iter().map(|v| fe_from_hex(v.as_str().unwrap())).collect();
    let mut coin = DefaultRandomCoin::<H>::new(&seed);
    let d = |key: &str| { let mut b = [0u8; 32]; b.copy_from_slice(&hex::decode(k["segment_proof"][key].as_str().unwrap()).unwrap()); <H as Hasher>::Digest::from(b) };
    coin.reseed(d("trace_root"));
    let _cc: Vec<BE> = (0..k["segment_proof"]["num_constraints"].as_u64().unwrap()).map(|_| coin.draw().unwrap()).collect();
    coin.reseed(d("constraint_root"));
    let z: BE = coin.draw().unwrap();
    assert_eq!(fe_hex(z), k["segment_proof"]["z"].as_str().unwrap());
}

// ---- GPU path (needs libzkl_b200.so + a B200) ----
#[test]
fn gpu_full_mode_yields_the_committed_bytes() {
    let s = segment();
    let ctx = zkl_b200_prover::GpuCtx::new(0);
    let np = (s.j["num_partitions"].as_u64().unwrap() as usize, s.j["hash_rate"].as_u64().unwrap() as usize);
    let proof = zkl_b200_prover::prove_segment_full(&ctx, s.trace.main_segment(), &s.air_pi, &s.opts, np).expect("zkl_prove_segment_pi");
    assert_eq!(proof.to_bytes(), s.proof_bytes);
}
#[test]
fn gpu_program_constants_equal_the_reference() {
    let s = segment();
    let p = zkl_b200_prover::zkl_program(&s.pi.program_id, &s.pi.program_commitment);
    let ps = poseidon::get_poseidon_suite(&s.pi.program_id);
    for i in 0..12 { for j in 0..12 { assert_eq!(u128::from(p.pose.mds[i][j]), ps.mds[i][j].as_int()); } }
    for r in 0..27 { for j in 0..12 { assert_eq!(u128::from(p.pose.rc[r][j]), ps.rc[r][j].as_int()); } }
    let w0 = utils::rom_weights_for_seed(utils::ROM_W_SEED_0);
    for k in 0..59 { assert_eq!(u128::from(p.rom_w_enc0[k]), w0[k].as_int()); }
    let fc = zk_lisp_proof_winterfell::commit::program_field_commitment(&s.pi.program_commitment);
    assert_eq!(u128::from(p.commitment_fe[0]), fc[0].as_int());
    assert_eq!(u128::from(p.commitment_fe[1]), fc[1].as_int());
}
