// host_air.inl — what `ZkLispAir::new` derives from the public inputs, built on the host in native code
// (included by zkl_b200.cu). Nothing here touches the device; every function is plain C++ over the f128 limb
// arithmetic of f128.cuh (its functions are __host__ __device__).
//
// Reference text followed:
//   poseidon/mod.rs:52-75,111-217   get_poseidon_suite (dom, Cauchy MDS from rejection-sampled points, round constants)
//   poseidon/mod.rs:186-197,219-251 ROM t=3 suite;   poseidon/mod.rs:421-440 ro_from_slices (blake3 -> lo + hi*2^64)
//   utils.rs:114-141 rom_weights_for_seed;            commit.rs:31-79 program_field_commitment
//   vm/layout.rs:183-314 Columns::for_config;         vm/air/mod.rs:114-318 ZkLispAir::new (degrees, declared assertions)
//   vm/air/*.rs push_degrees;                         vm/air/mod.rs:380-504, schedule.rs:42-140, rom.rs:122-147 assertions
//   lib.rs:116-160 AirPublicInputs::to_elements;      winter-air Context::to_elements [WF-mem]

// ---- BLAKE3, hash mode, inputs of at most one chunk (1024 bytes): every derivation input is a domain string plus < 64 bytes ----
static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static inline uint32_t b3_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void b3_g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
  s[a] = s[a] + s[b] + mx; s[d] = b3_rotr(s[d] ^ s[a], 16);
  s[c] = s[c] + s[d];      s[b] = b3_rotr(s[b] ^ s[c], 12);
  s[a] = s[a] + s[b] + my; s[d] = b3_rotr(s[d] ^ s[a], 8);
  s[c] = s[c] + s[d];      s[b] = b3_rotr(s[b] ^ s[c], 7);
}
static void b3_compress(uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint32_t flags) {
  static const int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
  uint32_t m[16], s[16];
  for (int i = 0; i < 16; i++) m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
  for (int i = 0; i < 8; i++) s[i] = cv[i];
  for (int i = 0; i < 4; i++) s[8 + i] = B3_IV[i];
  s[12] = 0; s[13] = 0; s[14] = block_len; s[15] = flags;   // chunk counter 0: single chunk
  for (int r = 0; r < 7; r++) {
    b3_g(s, 0, 4, 8, 12, m[0], m[1]);  b3_g(s, 1, 5, 9, 13, m[2], m[3]);
    b3_g(s, 2, 6, 10, 14, m[4], m[5]); b3_g(s, 3, 7, 11, 15, m[6], m[7]);
    b3_g(s, 0, 5, 10, 15, m[8], m[9]); b3_g(s, 1, 6, 11, 12, m[10], m[11]);
    b3_g(s, 2, 7, 8, 13, m[12], m[13]); b3_g(s, 3, 4, 9, 14, m[14], m[15]);
    if (r < 6) { uint32_t t[16]; for (int i = 0; i < 16; i++) t[i] = m[perm[i]]; memcpy(m, t, sizeof m); }
  }
  for (int i = 0; i < 8; i++) cv[i] = s[i] ^ s[i + 8];
}
static bool blake3_one_chunk(const uint8_t *data, size_t len, uint8_t out[32]) {
  if (len > 1024) return false;
  uint32_t cv[8]; memcpy(cv, B3_IV, sizeof cv);
  const size_t nblocks = std::max<size_t>(1, (len + 63) / 64);
  for (size_t b = 0; b < nblocks; b++) {
    uint8_t blk[64]; memset(blk, 0, 64);
    const size_t l = std::min<size_t>(64, len - 64 * b);
    if (len) memcpy(blk, data + 64 * b, l);
    uint32_t flags = (b == 0 ? 1u : 0u) | (b == nblocks - 1 ? (2u | 8u) : 0u);   // CHUNK_START, CHUNK_END | ROOT
    b3_compress(cv, blk, (uint32_t)l, flags);
  }
  for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(cv[i] >> (8 * k));
  return true;
}

// ---- parameter derivation ----
static inline zkl_fe fe_to_abi(const fe &x) { zkl_fe r; r.lo = (uint64_t)x.l[0] | ((uint64_t)x.l[1] << 32); r.hi = (uint64_t)x.l[2] | ((uint64_t)x.l[3] << 32); return r; }
static inline fe fe_from_le16(const uint8_t *b) { fe v; memcpy(v.l, b, 16); return fe_from_u128_mod(v); }   // 16 LE bytes as an integer, mod p
struct ro_input {
  uint8_t buf[160]; size_t len = 0;
  ro_input(const char *domain) { len = strlen(domain); memcpy(buf, domain, len); }
  ro_input &bytes(const void *p, size_t n) { memcpy(buf + len, p, n); len += n; return *this; }
  ro_input &u8(uint8_t v) { buf[len++] = v; return *this; }
  ro_input &u32le(uint32_t v) { for (int i = 0; i < 4; i++) buf[len++] = (uint8_t)(v >> (8 * i)); return *this; }
  fe value() const {   // poseidon/mod.rs:421-440: first 16 digest bytes, lo + hi * 2^64
    uint8_t d[32]; blake3_one_chunk(buf, len, d);
    return fe_from_le16(d);
  }
};
// poseidon/mod.rs:114-131: distinct non-zero points; the retry counter runs on across points
static void derive_points(const char *domain, const uint8_t sid[32], int n, fe *pts) {
  int have = 0; uint32_t ctr = 0;
  while (have < n) {
    fe cand = ro_input(domain).bytes(sid, 32).u8((uint8_t)have).u32le(ctr).value();
    bool ok = !fe_is_zero(cand);
    for (int i = 0; ok && i < have; i++) ok = !fe_eq(cand, pts[i]);
    if (ok) pts[have++] = cand; else ctr++;
  }
}
struct host_suite { fe dom[2], mds[12][12], rc[ZKL_ROUNDS][12]; };
static int derive_poseidon_suite(const uint8_t sid[32], host_suite &s) {
  s.dom[0] = ro_input("zkl/poseidon2/dom/c0").bytes(sid, 32).value();
  s.dom[1] = ro_input("zkl/poseidon2/dom/c1").bytes(sid, 32).value();
  fe x[12], y[12];
  derive_points("zkl/poseidon2/mds/x", sid, 12, x);
  derive_points("zkl/poseidon2/mds/y", sid, 12, y);
  for (uint32_t adj = 0;; adj++) {   // poseidon/mod.rs:146-170: re-draw y while some x_i + y_j vanishes
    bool bad = false;
    for (int i = 0; i < 12 && !bad; i++) for (int j = 0; j < 12; j++) if (fe_is_zero(fe_add(x[i], y[j]))) { bad = true; break; }
    if (!bad) break;
    if (adj > 1000000) return 1;
    for (int j = 0; j < 12; j++) {
      fe cand = ro_input("zkl/poseidon2/mds/y").bytes(sid, 32).u8((uint8_t)j).u32le(adj).value();
      y[j] = fe_is_zero(cand) ? fe_one() : cand;
    }
  }
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) s.mds[i][j] = fe_inv(fe_add(x[i], y[j]));
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int l = 0; l < 12; l++) s.rc[r][l] = ro_input("zkl/poseidon2/rc").bytes(sid, 32).u8((uint8_t)r).u8((uint8_t)l).value();
  return 0;
}
static void host_permute(const host_suite &s, fe st[12]) {   // poseidon/hasher.rs:173-190
  for (int r = 0; r < ZKL_ROUNDS; r++) {
    fe c3[12], nx[12];
    for (int i = 0; i < 12; i++) c3[i] = fe_cube(st[i]);
    for (int i = 0; i < 12; i++) {
      wide_acc w; wacc_zero(w);
      for (int k = 0; k < 12; k++) wacc_mad(w, c3[k], s.mds[i][k]);
      nx[i] = fe_add(wacc_reduce(w), s.rc[r][i]);
    }
    for (int i = 0; i < 12; i++) st[i] = nx[i];
  }
}
// commit.rs:31-79: two lanes of the permutation of [fe16(b[0..16]), fe16(b[16..32]), 0 x 8, dom0, dom1] under suite(b)
static int program_field_commitment_host(const uint8_t b[32], const host_suite *suite_of_b, fe out[2]) {
  host_suite tmp;
  if (!suite_of_b) { if (derive_poseidon_suite(b, tmp)) return 1; suite_of_b = &tmp; }
  fe st[12];
  for (int i = 0; i < 12; i++) st[i] = fe_zero();
  st[0] = fe_from_le16(b); st[1] = fe_from_le16(b + 16);
  st[10] = suite_of_b->dom[0]; st[11] = suite_of_b->dom[1];
  host_permute(*suite_of_b, st);
  out[0] = st[0]; out[1] = st[1];
  return 0;
}
static void suite_to_abi(const host_suite &s, zkl_suite *o) {
  for (int i = 0; i < 2; i++) o->dom[i] = fe_to_abi(s.dom[i]);
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) o->mds[i][j] = fe_to_abi(s.mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 12; j++) o->rc[r][j] = fe_to_abi(s.rc[r][j]);
}
static bool any_byte(const uint8_t *b, size_t n) { for (size_t i = 0; i < n; i++) if (b[i]) return true; return false; }

extern "C" int zkl_suite_derive(const uint8_t suite_id[32], zkl_suite *out) {
  host_suite s;
  if (derive_poseidon_suite(suite_id, s)) return 1;
  suite_to_abi(s, out);
  return 0;
}
extern "C" int zkl_program_derive(const uint8_t program_id[32], const uint8_t program_commitment[32], zkl_program *out) {
  memset(out, 0, sizeof *out);
  memcpy(out->program_id, program_id, 32); memcpy(out->program_commitment, program_commitment, 32);
  host_suite s;
  if (derive_poseidon_suite(program_id, s)) return 1;
  suite_to_abi(s, &out->pose);
  fe x[3], y[3];   // poseidon/mod.rs:219-251 (no re-draw loop for the 3x3 matrix)
  derive_points("zkl/rom3/mds/x", program_id, 3, x);
  derive_points("zkl/rom3/mds/y", program_id, 3, y);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out->rom.mds[i][j] = fe_to_abi(fe_inv(fe_add(x[i], y[j])));
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int l = 0; l < 3; l++) out->rom.rc[r][l] = fe_to_abi(ro_input("zkl/rom3/rc").bytes(program_id, 32).u8((uint8_t)r).u8((uint8_t)l).value());
  const fe three = fe_from_u64(3);   // utils.rs:114-141: 3^(seed+1) .. 3^(seed+59), seeds 17 and 1037 (utils.rs:27-28)
  fe w0 = fe_pow(three, 17 + 1), w1 = fe_pow(three, 1037 + 1);
  for (int k = 0; k < 59; k++) { out->rom_w_enc0[k] = fe_to_abi(w0); out->rom_w_enc1[k] = fe_to_abi(w1); w0 = fe_mul(w0, three); w1 = fe_mul(w1, three); }
  fe pf[2];
  if (any_byte(program_id, 32)) { if (program_field_commitment_host(program_id, &s, pf)) return 1; out->program_fe[0] = fe_to_abi(pf[0]); out->program_fe[1] = fe_to_abi(pf[1]); }
  if (any_byte(program_commitment, 32)) {
    if (program_field_commitment_host(program_commitment, nullptr, pf)) return 1;
    out->commitment_fe[0] = fe_to_abi(pf[0]); out->commitment_fe[1] = fe_to_abi(pf[1]);
  }
  return 0;
}

// ---- column map: vm/layout.rs:183-314 Columns::for_config (only ram / merkle / rom change the layout) ----
static zkl_columns columns_for_config(bool ram, bool merkle, bool rom) {
  zkl_columns c; memset(&c, 0, sizeof c);
  uint32_t cur = 0;
  c.lanes_start = cur; cur += 12;
  c.g_map = cur; c.g_final = cur + 1; c.g_r_start = cur + 2; cur = c.g_r_start + ZKL_ROUNDS;
  c.mask = cur; cur += 1;
  c.r_start = cur; cur += ZKL_NR;
  c.op_start = cur; cur += ZKL_NUM_OPS;
  c.sel_dst0_start = cur; c.sel_a_start = cur + ZKL_NR; c.sel_b_start = cur + 2 * ZKL_NR; c.sel_c_start = cur + 3 * ZKL_NR; c.sel_dst1_start = cur + 4 * ZKL_NR;
  cur += 5 * ZKL_NR;
  c.sel_s_bits_start = cur; c.sel_s_active_start = cur + 10 * 3; cur = c.sel_s_active_start + 10;
  c.imm = cur; c.eq_inv = cur + 1; cur += 2;
  c.ram_start = cur; if (ram) cur += 8;
  c.merkle_start = cur; if (merkle) cur += 7;
  c.pi_prog = cur; cur += 1;
  c.pc = cur; cur += 1;
  c.rom_op_start = cur; if (rom) cur += 17;
  c.pose_active = cur; cur += 1;
  c.gadget_b_start = cur; cur += 32;
  c.rom_s_start = cur; if (rom) cur += 3;
  c.width = rom ? cur : c.pc + 1;   // layout.rs:310-313
  return c;
}

// ---- constraint degrees: only their count and two maxima are needed (AirContext::new [WF-mem]) ----
struct degree_stats {
  uint32_t count = 0, ce_blowup = 2; uint64_t highest = 0; size_t n;
  void push(uint32_t base, uint32_t cycles, uint32_t times = 1) {
    if (!times) return;
    count += times;
    uint32_t p2 = 1; while (p2 < base + cycles) p2 <<= 1;
    ce_blowup = std::max(ce_blowup, p2);
    highest = std::max<uint64_t>(highest, (uint64_t)base * (n - 1) + (uint64_t)cycles * (n / 32) * 31);   // every cycle has length 32
  }
};
enum { FM_POSEIDON = 1u << 0, FM_VM = 1u << 1, FM_VM_EXPECT = 1u << 4, FM_SPONGE = 1u << 5, FM_MERKLE = 1u << 6, FM_RAM = 1u << 7 };
enum { VMU_ASSERT = 0, VMU_ASSERT_BIT, VMU_ASSERT_RANGE, VMU_DIVMOD, VMU_MULWIDE, VMU_DIV128, VMU_EQ, VMU_SPONGE, VMU_RAM_DELTA_CLK };
static degree_stats transition_degree_stats(uint64_t fm, bool rom_enabled, uint32_t usage, uint32_t ram_bits, size_t n) {
  degree_stats d; d.n = n;
  const bool sponge_used = (usage >> VMU_SPONGE) & 1;
  auto bit = [&](int b) { return (usage >> b) & 1u; };
  if (fm & FM_POSEIDON) {   // poseidon.rs:26-65
    d.push(4, 1, ZKL_ROUNDS * 12); d.push(1, 1, 12);
    if ((fm & FM_VM) && (fm & FM_SPONGE) && sponge_used) { d.push(6, 1, 2); d.push(3, 1, 8); }
  }
  if (fm & FM_VM) {
    d.push(2, 1, 5 * ZKL_NR); d.push(1, 1, 5); d.push(2, 1, ZKL_NR);   // ctrl.rs:29-112
    if ((fm & FM_SPONGE) && sponge_used) d.push(2, 1, 10 * 4);
    d.push(2, 1, 1 + 17 + 1 + 17); d.push(1, 1, 2);
    d.push(1, 1, ZKL_NR); d.push(7, 1, ZKL_NR);                       // alu.rs:30-105
    d.push(5, 1, 2 * bit(VMU_EQ)); d.push(5, 1, 2 * bit(VMU_DIVMOD)); d.push(5, 1, bit(VMU_ASSERT)); d.push(5, 1, bit(VMU_ASSERT_BIT));
    d.push(5, 1, 33 * bit(VMU_ASSERT_RANGE)); d.push(5, 1, bit(VMU_MULWIDE)); d.push(5, 1, 2 * bit(VMU_DIV128));
  }
  if (fm & FM_RAM) {   // ram.rs:26-79
    d.push(4, 1); d.push(2, 0); d.push(5, 0); d.push(3, 0); d.push(6, 0); d.push(5, 0);
    if (bit(VMU_RAM_DELTA_CLK)) d.push(5, 0, (uint32_t)__builtin_popcount(ram_bits) + 1);
    d.push(2, 0);
  }
  if (fm & FM_MERKLE) { d.push(3, 1, 3); d.push(2, 1); d.push(3, 1, 3); }   // merkle.rs:28-64
  if (rom_enabled) { d.push(3, 1, ZKL_ROUNDS * 3); d.push(1, 1, 3 + 2); }   // rom.rs:27-55
  if (!d.count) d.push(1, 0);                                               // vm/air/mod.rs:302-306
  return d;
}

// ---- the whole AIR instance ----
// The assertion list in closed form: `tmpl` holds the assertions of ONE schedule level (step = position inside the level, 141 of them
// for ZkLispAir), repeated for every one of `levels` levels of `period` rows; `extras` the few public-input-dependent single
// assertions at absolute steps. The device expands it in Winterfell's order (api_stark.inl: upload_air) — a 2^22-row segment declares
// 18.5 M assertions, which the host would otherwise build, order and push through PCIe for every proof.
struct sched_assertions {
  uint32_t period = 32, levels = 0;
  std::vector<zkl_assertion> tmpl, extras;
  size_t count() const { return (size_t)levels * tmpl.size() + extras.size(); }
};
struct built_air {
  zkl_air_params p;
  std::vector<zkl_assertion> as;   // the explicit list (expand = true), get_assertions order
  sched_assertions sched;          // the same list in closed form
  std::vector<zkl_fe> seed;
  std::string err;
};
static inline zkl_fe abi_u64(uint64_t v) { zkl_fe r; r.lo = v; r.hi = 0; return r; }
static inline bool abi_is_zero(const zkl_fe &v) { return (v.lo | v.hi) == 0; }
static inline zkl_fe be_from_le8_abi(const uint8_t b[32]) { return fe_to_abi(fe_from_le16(b)); }   // utils.rs:346-355

static int build_air_host(const zkl_program *prog, const zkl_public_inputs *pi, size_t width, size_t n, const zkl_options *o, built_air &out,
                          bool expand = true) {
  if (n < 32 || (n & (n - 1))) { out.err = "trace length must be a power of two >= 32 (one schedule level)"; return 1; }
  if (pi->num_main_slots > ZKL_MAX_MAIN_SLOTS) { out.err = "too many main_args slots"; return 1; }
  const uint64_t fm = pi->segment_feature_mask ? pi->segment_feature_mask : pi->feature_mask;
  const bool rom_layout = any_byte(prog->program_id, 32), has_commit = any_byte(prog->program_commitment, 32);
  // vm/air/mod.rs:236-238 pushes the ROM degrees under program_id != 0 while the constraints are gated by program_commitment
  // (mod.rs:355, ctrl.rs:332); the CLI always sets both (zk-lisp-proof/src/pi.rs:96-105)
  if (rom_layout != has_commit) { out.err = "program_id and program_commitment must be both zero or both non-zero"; return 1; }
  const uint32_t baseline_w = columns_for_config(true, true, true).width;
  const zkl_columns cols = width < baseline_w ? columns_for_config(fm & FM_RAM, fm & FM_MERKLE, rom_layout) : columns_for_config(true, true, rom_layout);
  if (cols.width != width) { out.err = "trace width " + std::to_string(width) + " does not match layout width " + std::to_string(cols.width); return 1; }
  zkl_air_params &p = out.p; memset(&p, 0, sizeof p);
  p.cols = cols;
  p.f_poseidon = !!(fm & FM_POSEIDON); p.f_vm = !!(fm & FM_VM); p.f_sponge = !!(fm & FM_SPONGE); p.f_merkle = !!(fm & FM_MERKLE); p.f_ram = !!(fm & FM_RAM);
  p.rom_enabled = has_commit; p.vm_usage_mask = pi->vm_usage_mask; p.ram_delta_clk_bits = pi->ram_delta_clk_bits;
  const degree_stats d = transition_degree_stats(fm, has_commit, pi->vm_usage_mask, pi->ram_delta_clk_bits, n);
  p.num_transition = d.count; p.ce_blowup = d.ce_blowup;
  p.num_composition_cols = (uint32_t)std::max<uint64_t>((d.highest - (n - 1) + n - 1) / n, 1);   // divisor degree n - 1 (one exemption)
  p.pose = prog->pose; p.rom = prog->rom;
  memcpy(p.rom_w_enc0, prog->rom_w_enc0, sizeof p.rom_w_enc0); memcpy(p.rom_w_enc1, prog->rom_w_enc1, sizeof p.rom_w_enc1);
  p.program_fe0 = prog->program_fe[0]; p.merkle_root_fe = be_from_le8_abi(pi->merkle_root);

  // assertions (vm/air/mod.rs:380-504), de-duplicated by (column, step) keeping the first. The schedule assertions of a level
  // (schedule.rs:42-140) are the same for every level: they go into the template once; everything that depends on the public inputs
  // is an extra. An extra that names a cell the schedule already asserts is dropped (the first one wins), like a repeated extra.
  const uint32_t last = (uint32_t)n - 1, levels = (uint32_t)(n / 32);
  const bool vm = fm & FM_VM, pc0 = abi_is_zero(pi->pc_init);
  sched_assertions &sa = out.sched;
  sa.period = 32; sa.levels = levels; sa.tmpl.clear(); sa.extras.clear();
  const zkl_fe one = abi_u64(1), zero = abi_u64(0);
  auto put_t = [&](uint32_t col, uint32_t pos, const zkl_fe &v) {
    for (const auto &x : sa.tmpl) if (x.column == col && x.step == pos) return;
    sa.tmpl.push_back(zkl_assertion{col, pos, v});
  };
  {
    const uint32_t row_map = 0, row_final = 1 + ZKL_ROUNDS;
    put_t(cols.lanes_start + 10, row_map, prog->pose.dom[0]);
    put_t(cols.lanes_start + 11, row_map, prog->pose.dom[1]);
    put_t(cols.g_map, row_map, one);
    put_t(cols.g_final, row_final, one);
    for (uint32_t j = 0; j < ZKL_ROUNDS; j++) put_t(cols.g_r_start + j, 1 + j, one);
    put_t(cols.g_final, row_map, zero);
    for (uint32_t j = 0; j < ZKL_ROUNDS; j++) put_t(cols.g_r_start + j, row_map, zero);
    put_t(cols.g_map, row_final, zero);
    for (uint32_t j = 0; j < ZKL_ROUNDS; j++) put_t(cols.g_r_start + j, row_final, zero);
    for (uint32_t j = 0; j < ZKL_ROUNDS; j++) { put_t(cols.g_map, 1 + j, zero); put_t(cols.g_final, 1 + j, zero); }
  }
  auto put_e = [&](uint32_t col, uint32_t step, const zkl_fe &v) {
    for (const auto &x : sa.tmpl) if (x.column == col && x.step == step % 32) return;
    for (const auto &x : sa.extras) if (x.column == col && x.step == step) return;
    sa.extras.push_back(zkl_assertion{col, step, v});
  };
  if (vm) {   // level 0 of the schedule loop (schedule.rs), then mod.rs:388-446
    if (pc0 && has_commit) put_e(cols.pi_prog, 0, be_from_le8_abi(prog->program_commitment));
    put_e(cols.pc, 0, pi->pc_init);
    if (fm & FM_VM_EXPECT) put_e(cols.r_start + std::min<uint32_t>(pi->vm_out_reg, ZKL_NR - 1), std::min<uint32_t>(pi->vm_out_row, last), be_from_le8_abi(pi->vm_expected_bytes));
    if (pc0 && pi->num_main_slots) {
      if (pi->num_main_slots > ZKL_NR) { out.err = "main_args must fit into NR registers"; return 1; }
      for (uint32_t j = 0; j < pi->num_main_slots; j++) put_e(cols.r_start + ZKL_NR - pi->num_main_slots + j, 0, pi->main_slots[j]);
    }
  }
  if (has_commit) {   // rom.rs:122-147
    for (uint32_t i = 0; i < 3; i++) put_e(cols.rom_s_start + i, 0, pi->rom_s_in[i]);
    for (uint32_t i = 0; i < 3; i++) put_e(cols.rom_s_start + i, last, pi->rom_s_out[i]);
  }
  // the explicit list in get_assertions order: every level's schedule assertions (the level-0 extras right behind level 0's), then the rest
  std::vector<zkl_assertion> &as = out.as;
  as.clear();
  if (expand) {
    as.reserve(sa.count());
    size_t lvl0_extras = vm ? (pc0 && has_commit ? 2 : 1) : 0;   // pi_prog / pc sit inside the level loop of the reference
    lvl0_extras = std::min(lvl0_extras, sa.extras.size());
    for (uint32_t lvl = 0; lvl < levels; lvl++) {
      for (const auto &x : sa.tmpl) as.push_back(zkl_assertion{x.column, lvl * 32 + x.step, x.value});
      if (lvl == 0) for (size_t e = 0; e < lvl0_extras; e++) as.push_back(sa.extras[e]);
    }
    for (size_t e = lvl0_extras; e < sa.extras.size(); e++) as.push_back(sa.extras[e]);
  }
  // the count ZkLispAir::new declares to AirContext (vm/air/mod.rs:240-290); Winterfell panics when get_assertions disagrees
  uint64_t declared = (uint64_t)(2 + ZKL_ROUNDS) * levels + (uint64_t)(4 * ZKL_ROUNDS + 2) * levels + 2 * (uint64_t)levels;
  if (vm) declared += 1 + (pc0 && has_commit ? 1 : 0) + (pc0 ? pi->num_main_slots : 0) + ((fm & FM_VM_EXPECT) ? 1 : 0);
  if (rom_layout) declared += 6;
  declared = std::max<uint64_t>(declared, 1);
  if (declared != sa.count()) { out.err = "declared " + std::to_string(declared) + " assertions but get_assertions yields " + std::to_string(sa.count()); return 1; }
  if (sa.count() >> 32) { out.err = "too many assertions"; return 1; }
  p.num_assertions = (uint32_t)sa.count(); p.assertions = expand ? as.data() : nullptr;

  // public-coin seed: Context::to_elements() [WF-mem] ++ AirPublicInputs::to_elements() (lib.rs:116-160)
  std::vector<zkl_fe> &s = out.seed; s.clear();
  s.push_back(abi_u64(((uint64_t)width << 8) | 0)); s.push_back(abi_u64(n));
  s.push_back(abi_u64(0xffffd30000000001ull)); s.push_back(abi_u64(0xffffffffffffffffull));   // modulus, LE halves
  s.push_back(abi_u64((uint64_t)p.num_transition + p.num_assertions));
  s.push_back(abi_u64((((uint64_t)1 /* FieldExtension::None */ << 8 | o->fri_folding) << 8) | o->fri_remainder_max_degree));
  s.push_back(abi_u64(o->grinding)); s.push_back(abi_u64(o->blowup)); s.push_back(abi_u64(o->num_queries));
  s.push_back(abi_u64(pi->feature_mask));
  s.push_back(be_from_le8_abi(prog->program_commitment));
  s.push_back(be_from_le8_abi(pi->merkle_root));
  s.push_back(prog->commitment_fe[0]); s.push_back(prog->commitment_fe[1]);   // zero when program_commitment is zero
  for (uint32_t j = 0; j < pi->num_main_slots; j++) s.push_back(pi->main_slots[j]);
  s.push_back(pi->pc_init);
  s.push_back(pi->ram_gp_unsorted_in); s.push_back(pi->ram_gp_unsorted_out); s.push_back(pi->ram_gp_sorted_in); s.push_back(pi->ram_gp_sorted_out);
  for (int i = 0; i < 3; i++) s.push_back(pi->rom_s_in[i]);
  for (int i = 0; i < 3; i++) s.push_back(pi->rom_s_out[i]);
  s.push_back(abi_u64(pi->vm_usage_mask)); s.push_back(abi_u64(pi->ram_delta_clk_bits));
  return 0;
}

extern "C" int zkl_air_build(const zkl_program *program, const zkl_public_inputs *pi, size_t width, size_t n, const zkl_options *opts,
                             zkl_air_params *out_params, zkl_assertion *out_assertions, size_t *num_assertions, zkl_fe *out_seed,
                             size_t *num_seed, char *err, size_t err_cap) {
  built_air b;
  auto fail = [&](const std::string &m) { if (err && err_cap) { snprintf(err, err_cap, "%s", m.c_str()); } return 1; };
  if (build_air_host(program, pi, width, n, opts, b)) return fail(b.err);
  const bool fits = b.as.size() <= *num_assertions && b.seed.size() <= *num_seed;
  *num_assertions = b.as.size(); *num_seed = b.seed.size();
  if (!fits) return fail("zkl_air_build: output buffers too small");
  memcpy(out_assertions, b.as.data(), b.as.size() * sizeof(zkl_assertion));
  memcpy(out_seed, b.seed.data(), b.seed.size() * sizeof(zkl_fe));
  *out_params = b.p; out_params->assertions = out_assertions;
  return 0;
}

// ---- ZlAggAir (agg/air.rs:22-332): 31 columns, 24 transition constraints, 5 assertions, one periodic column of cycle n ----
// Degrees (agg/air.rs:69-94): constraint 1 has base degree 2 and one cycle-n periodic factor, constraint 11 base 1 and one
// periodic factor, the rest are linear. AirContext [WF-mem]: ce_blowup = next_pow2(2 + 1) = 4; a cycle-n periodic column has
// degree n - 1, so the highest evaluation degree is 3 (n - 1) and the composition polynomial has (3(n-1) - (n-1) + n - 1) / n columns.
extern "C" int zkl_agg_air_params(size_t n, uint64_t v_units_total, uint32_t children_count, zkl_air_params *p, zkl_assertion *as) {
  if (n < 8 || (n & (n - 1)) || n > ((size_t)1 << 24)) return 1;   // agg/trace.rs:35,396-399: max(children, 8) rounded up to a power of two
  memset(p, 0, sizeof *p);
  p->air_kind = ZKL_AIR_AGG;
  p->cols.width = ZKL_AGG_WIDTH;
  p->num_transition = ZKL_AGG_TRANSITIONS;
  p->ce_blowup = 4;
  p->num_composition_cols = (uint32_t)std::max<uint64_t>((3 * (uint64_t)(n - 1) - (n - 1) + n - 1) / n, 1);
  const uint32_t last = (uint32_t)n - 1, ok = 0, v_units_acc = 22, child_count_acc = 24;   // agg/layout.rs:97-180
  as[0] = zkl_assertion{ok, 0, abi_u64(0)};
  as[1] = zkl_assertion{v_units_acc, 0, abi_u64(0)};
  as[2] = zkl_assertion{v_units_acc, last, abi_u64(v_units_total)};
  as[3] = zkl_assertion{child_count_acc, 0, abi_u64(0)};
  as[4] = zkl_assertion{child_count_acc, last, abi_u64(children_count)};
  p->num_assertions = 5; p->assertions = as;
  return 0;
}
