// api_replay.inl — child-proof replay for a batch of step proofs, native host side over the device hasher.
//
// What the aggregation layer of the reference does per child on the CPU hasher — zk-lisp-proof-winterfell/src/agg/fs.rs:37-245
// (Fiat-Shamir replay: coin seed from the proof context + AIR public inputs, reseeds with the commitments, OOD point, DEEP
// coefficients, FRI alphas, grinding check, query positions) and src/agg/child.rs:1025-1099 (row leaves incl. partitions, Merkle path
// folding, FRI position folding) — is done here for ALL children of a batch in lockstep: every transcript step is one device batch
// over the children, the leaves of all opened rows / FRI pairs are three hashing batches and every tree of every child is
// authenticated by ONE merge plan (zkl_merge_plan). The proof bytes are the ones zkl_prove_segment wrote (byte_writer above;
// winter-air 0.13 Proof::to_bytes [WF-mem]). zk-lisp_b200/replay.py is the same logic in Python and the test reference for this file.

struct rp_span { const uint8_t *p = nullptr; size_t n = 0; };
struct rp_reader {
  const uint8_t *p; size_t n, i = 0; bool ok = true;
  rp_reader(const uint8_t *p_, size_t n_) : p(p_), n(n_) {}
  rp_span take(size_t k) { rp_span s; if (!ok || k > n - i) { ok = false; return s; } s.p = p + i; s.n = k; i += k; return s; }
  uint64_t uint(int k) { rp_span s = take((size_t)k); uint64_t v = 0; if (ok) for (int j = k - 1; j >= 0; j--) v = (v << 8) | s.p[j]; return v; }
  uint64_t usize() {  // winter-utils vint64: trailing zeros of the first byte give the length
    if (!ok || i >= n) { ok = false; return 0; }
    uint8_t first = p[i];
    if (first == 0) { i++; return uint(8); }
    int len = __builtin_ctz(first) + 1;
    uint64_t v = uint(len);
    return ok ? v >> len : 0;
  }
};
struct rp_view {
  uint32_t W = 0, aux = 0, logn = 0, q = 0, blowup = 0, grind = 0, ext = 0, folding = 0, rem_deg = 0, parts = 0, rate = 0, unique = 0;
  uint64_t num_constraints = 0, nonce = 0;
  rp_span commitments, tq_values, tq_paths, cq_values, cq_paths, ood_trace, ood_cons, remainder;
  std::vector<std::pair<rp_span, rp_span>> layers;
  size_t n() const { return (size_t)1 << logn; }
  size_t N() const { return n() * blowup; }
  size_t num_fri_layers() const { size_t d = N(), max_rem = (size_t)(rem_deg + 1) * blowup, r = 0; while (d > max_rem) { d /= folding; r++; } return r; }
};
static bool rp_parse(const uint8_t *p, size_t len, rp_view &v) {
  rp_reader r(p, len);
  v.W = (uint32_t)r.uint(1); v.aux = (uint32_t)r.uint(1); v.logn = (uint32_t)r.uint(1);
  r.take((size_t)r.uint(2));                                   // trace meta
  r.take((size_t)r.uint(1));                                   // field modulus bytes
  v.num_constraints = r.usize();
  v.q = (uint32_t)r.uint(1); v.blowup = (uint32_t)r.uint(1); v.grind = (uint32_t)r.uint(1); v.ext = (uint32_t)r.uint(1);
  v.folding = (uint32_t)r.uint(1); v.rem_deg = (uint32_t)r.uint(1);
  r.take(2);                                                   // batching methods
  v.parts = (uint32_t)r.uint(1); v.rate = (uint32_t)r.uint(1);
  v.unique = (uint32_t)r.uint(1);
  if (!r.ok || v.W == 0 || v.logn >= 40 || v.blowup == 0 || v.folding != 2 || v.parts == 0) return false;
  v.commitments = r.take((size_t)r.uint(2));
  v.tq_values = r.take((size_t)r.usize()); v.tq_paths = r.take((size_t)r.usize());
  v.cq_values = r.take((size_t)r.usize()); v.cq_paths = r.take((size_t)r.usize());
  v.ood_trace = r.take((size_t)r.uint(2)); v.ood_cons = r.take((size_t)r.uint(2));
  size_t nl = (size_t)r.uint(1);
  for (size_t d = 0; d < nl && r.ok; d++) { rp_span a = r.take((size_t)r.uint(4)); rp_span b = r.take((size_t)r.uint(4)); v.layers.push_back({a, b}); }
  v.remainder = r.take((size_t)r.uint(2));
  r.uint(1);                                                   // FRI proof partitions
  v.nonce = r.uint(8);
  return r.ok && r.i == r.n;
}
static bool rp_read_batch_proof(rp_span s, int &depth, std::vector<std::vector<const uint8_t *>> &lists) {
  rp_reader r(s.p, s.n);
  depth = (int)r.uint(1);
  uint64_t k = r.usize();
  if (!r.ok || k > (1u << 20)) return false;
  lists.assign((size_t)k, {});
  for (auto &l : lists) {
    uint64_t m = r.usize();
    if (!r.ok || m > 64) return false;
    for (uint64_t j = 0; j < m; j++) { rp_span d = r.take(32); if (!r.ok) return false; l.push_back(d.p); }
  }
  return r.ok && r.i == r.n;
}

// merge schedule of many batch openings (replay.py: MergePlan)
struct rp_node { int level; uint32_t idx; };   // level -1: known digest
struct rp_plan {
  std::vector<const uint8_t *> known;
  std::vector<std::vector<std::pair<rp_node, rp_node>>> levels;
  struct root_t { rp_node node; const uint8_t *expected; size_t child; int status; };
  std::vector<root_t> roots;
  // BatchMerkleProof::get_root [WF-mem] recorded as merges; returns false (plan untouched) when the opening is malformed
  bool add_opening(int depth, const std::vector<std::vector<const uint8_t *>> &lists, const std::vector<const uint8_t *> &leaves,
                   const std::vector<uint32_t> &positions, const uint8_t *expected, size_t child, int status) {
    if (depth < 1 || depth > 40 || leaves.size() != positions.size() || positions.empty()) return false;
    const uint64_t n = (uint64_t)1 << depth;
    std::vector<uint64_t> norm;
    for (uint32_t p : positions) { if (p >= n) return false; norm.push_back(p & ~(uint64_t)1); }
    std::sort(norm.begin(), norm.end()); norm.erase(std::unique(norm.begin(), norm.end()), norm.end());
    if (norm.size() != lists.size()) return false;
    std::vector<const uint8_t *> lk;                                       // local known digests
    std::vector<std::vector<std::pair<rp_node, rp_node>>> ll((size_t)depth);   // local levels
    std::map<uint64_t, rp_node> opened;
    for (size_t i = 0; i < positions.size(); i++) { opened[positions[i]] = rp_node{-1, (uint32_t)lk.size()}; lk.push_back(leaves[i]); }
    std::vector<size_t> ptr(lists.size(), 0);
    bool bad = false;
    auto take = [&](size_t s) { if (ptr[s] >= lists[s].size()) { bad = true; return rp_node{-1, 0}; } lk.push_back(lists[s][ptr[s]++]); return rp_node{-1, (uint32_t)(lk.size() - 1)}; };
    auto merge = [&](int lvl, rp_node a, rp_node b) { ll[(size_t)lvl].push_back({a, b}); return rp_node{lvl, (uint32_t)(ll[(size_t)lvl].size() - 1)}; };
    std::vector<uint64_t> cur; std::vector<rp_node> vals;
    for (size_t i = 0; i < norm.size(); i++) {
      auto lo_it = opened.find(norm[i]), hi_it = opened.find(norm[i] + 1);
      rp_node lo = lo_it != opened.end() ? lo_it->second : take(i);
      rp_node hi = hi_it != opened.end() ? hi_it->second : take(i);
      cur.push_back((norm[i] + n) >> 1); vals.push_back(merge(0, lo, hi));
    }
    for (int lvl = 1; lvl < depth && !bad; lvl++) {
      std::vector<uint64_t> ncur; std::vector<rp_node> nvals;
      for (size_t i = 0; i < cur.size(); i++) {
        const size_t used = i; const uint64_t sib = cur[i] ^ 1;
        rp_node me = vals[i], sv;
        if (i + 1 < cur.size() && cur[i + 1] == sib) { sv = vals[i + 1]; i++; }
        else sv = take(used);
        nvals.push_back((cur[used] & 1) ? merge(lvl, sv, me) : merge(lvl, me, sv));
        ncur.push_back(sib >> 1);
      }
      cur.swap(ncur); vals.swap(nvals);
    }
    if (bad || cur.size() != 1 || cur[0] != 1) return false;
    for (size_t s = 0; s < lists.size(); s++) if (ptr[s] != lists[s].size()) return false;
    // append with remapped indices
    const uint32_t kbase = (uint32_t)known.size();
    if (levels.size() < (size_t)depth) levels.resize((size_t)depth);
    std::vector<uint32_t> lbase((size_t)depth);
    for (int l = 0; l < depth; l++) lbase[(size_t)l] = (uint32_t)levels[(size_t)l].size();
    auto remap = [&](rp_node x) { return x.level < 0 ? rp_node{-1, x.idx + kbase} : rp_node{x.level, x.idx + lbase[(size_t)x.level]}; };
    known.insert(known.end(), lk.begin(), lk.end());
    for (int l = 0; l < depth; l++) for (auto &m : ll[(size_t)l]) levels[(size_t)l].push_back({remap(m.first), remap(m.second)});
    roots.push_back({remap(vals[0]), expected, child, status});
    return true;
  }
};

#ifndef ZKL_REPLAY_HOST_ONLY   // tests/host_replay.cc compiles the parser and the plan builder above with g++
// poseidon_hash_two_lanes (poseidon/mod.rs:255-291) under an ARBITRARY suite, on the host: out[i] = lane 0 of the permutation of
// [left[i], right[i], 0 x 8, dom0, dom1]. The step digest (proof/digest.rs: 6 calls) and the children root (agg/child.rs:853-895:
// 2k-1 calls) use the PROGRAM's suite, not the commitment hasher's that lives in device constant memory; a handful of calls per
// proof — the same limb arithmetic as the kernels (f128.cuh compiles for the host), no device round trip, no context.
extern "C" int zkl_hash_two_lanes_host(const zkl_suite *suite, const zkl_fe *left, const zkl_fe *right, size_t count, zkl_fe *out) {
  if (!suite || (count && (!left || !right || !out))) return 1;
  fe mds[12][12], rc[ZKL_ROUNDS][12], dom[2];
  for (int i = 0; i < 2; i++) dom[i] = fe_from_abi(suite->dom[i]);
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) mds[i][j] = fe_from_abi(suite->mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 12; j++) rc[r][j] = fe_from_abi(suite->rc[r][j]);
  for (size_t q = 0; q < count; q++) {
    fe st[12];
    st[0] = fe_from_abi(left[q]); st[1] = fe_from_abi(right[q]);
    for (int i = 2; i < 10; i++) st[i] = fe_zero();
    st[10] = dom[0]; st[11] = dom[1];
    for (int r = 0; r < ZKL_ROUNDS; r++) {
      fe c3[12], y[12];
      for (int i = 0; i < 12; i++) c3[i] = fe_cube(st[i]);
      for (int i = 0; i < 12; i++) {
        wide_acc w; wacc_zero(w);
        for (int k = 0; k < 12; k++) wacc_mad(w, c3[k], mds[i][k]);
        y[i] = fe_add(wacc_reduce(w), rc[r][i]);
      }
      for (int i = 0; i < 12; i++) st[i] = y[i];
    }
    out[q].lo = (uint64_t)st[0].l[0] | ((uint64_t)st[0].l[1] << 32);
    out[q].hi = (uint64_t)st[0].l[2] | ((uint64_t)st[0].l[3] << 32);
  }
  return 0;
}

// Proof-of-work search on its own (ProverChannel::grind [WF-mem]; check restated at agg/fs.rs:202-207): the smallest nonce >= 1 with
// trailing_zeros(u64_le(merge_with_int(seed, nonce)[0..8])) >= bits. The aggregation proof of the reference stays on the CPU prover,
// but its grinding is 2^bits sponge calls on the slow CPU hasher — the one step of that proof worth sending to the device.
extern "C" int zkl_grind(zkl_ctx *ctx, const uint8_t seed[32], uint32_t bits, uint64_t *out_nonce) {
  ZKL_TRY(need_suite(ctx));
  if (bits > 40) { ctx->set_error("grinding factor too large"); return 1; }
  if (bits == 0) { *out_nonce = 0; return 0; }
  fe sd = fold_digest_host(seed);
  dev_buf<fe> coin; dev_buf<unsigned long long> nonce;
  ZKL_CHECK(ctx, coin.alloc(1)); ZKL_CHECK(ctx, nonce.alloc(1));
  ZKL_CHECK(ctx, cudaMemcpyAsync(coin.p, &sd, sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemsetAsync(nonce.p, 0xFF, 8, ctx->stream));
  unsigned long long nonce_h = ~0ull;
  for (uint64_t base = 1;;) {
    const uint64_t target = std::min<uint64_t>((uint64_t)16 << bits, (uint64_t)1 << 30);   // 16x the expected nonce per submission
    if (ctx->hash_path == 0) { LAUNCH_TC_TILES(ctx, k_tc_grind, target, coin.p, base, (uint32_t)target, bits, nonce.p); base += target; }
    else {
      const uint32_t batch = 1u << 15;
      int nb = (int)std::min<uint64_t>(std::max<uint64_t>(target / batch, 1), 512);
      for (int b = 0; b < nb; b++, base += batch)
        LAUNCH(ctx, k_grind, ceil_div((size_t)batch * ZKL_COOP_GROUP, ZKL_COOP_THREADS), ZKL_COOP_THREADS, 0, coin.p, base, batch, bits, nonce.p);
    }
    ZKL_CHECK(ctx, cudaMemcpyAsync(&nonce_h, nonce.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (nonce_h != ~0ull) break;
  }
  *out_nonce = nonce_h;
  return 0;
}

enum { RP_OK = 0, RP_MALFORMED = 1, RP_COMMITMENTS = 2, RP_OOD = 3, RP_FRI_DEGREE = 4, RP_POW = 5, RP_UNIQUE = 6, RP_TRACE_OPENING = 10,
       RP_CONSTRAINT_OPENING = 11, RP_FRI_OPENING = 12, RP_REMAINDER = 13, RP_BATCH_PROOF = 14 };
extern "C" const char *zkl_replay_status_name(int status) {
  switch (status) {
    case RP_OK: return "ok";
    case RP_MALFORMED: return "failed to decode inner Winterfell proof in step proof";
    case RP_COMMITMENTS: return "invalid Winterfell commitments layout in FS replay";
    case RP_OOD: return "invalid OOD frame layout in FS replay";
    case RP_FRI_DEGREE: return "FRI degree truncation detected while replaying FS challenges";
    case RP_POW: return "pow_nonce does not satisfy grinding factor in FS replay";
    case RP_UNIQUE: return "unique query positions count does not match proof.num_unique_queries";
    case RP_TRACE_OPENING: return "trace openings do not match the trace commitment";
    case RP_CONSTRAINT_OPENING: return "constraint openings do not match the constraint commitment";
    case RP_FRI_OPENING: return "FRI layer openings do not match the layer commitment";
    case RP_REMAINDER: return "FRI remainder does not match its commitment";
    case RP_BATCH_PROOF: return "malformed batch Merkle opening";
    default: return "unknown replay status";
  }
}

extern "C" int zkl_replay_children(zkl_ctx *ctx, const zkl_child_proof *children, size_t count, int verify_openings,
                                   zkl_child_challenges *out, zkl_fe *out_deep, size_t deep_stride) {
  ZKL_TRY(need_suite(ctx));
  if (count == 0) return 0;
  typedef std::array<uint8_t, 32> dig;
  std::vector<rp_view> v(count);
  std::vector<int> st(count, RP_OK);
  auto live = [&](size_t i) { return st[i] == RP_OK; };
  for (size_t i = 0; i < count; i++) {
    memset(&out[i], 0, sizeof out[i]);
    if (!rp_parse(children[i].proof, children[i].proof_len, v[i]) || v[i].aux != 0 || v[i].q == 0) { st[i] = RP_MALFORMED; continue; }
    const size_t nl = v[i].num_fri_layers();
    if (v[i].commitments.n != 32 * (2 + nl + 1) || nl + 1 > 40) { st[i] = RP_COMMITMENTS; continue; }
    if (v[i].ood_trace.n != 32u * v[i].W || v[i].ood_cons.n == 0 || v[i].ood_cons.n % 32) { st[i] = RP_OOD; continue; }
    if (v[i].W + v[i].ood_cons.n / 32 > deep_stride) { ctx->set_error("deep coefficient buffer too small"); return 1; }
  }
  auto root_of = [&](size_t i, size_t k) { return v[i].commitments.p + 32 * k; };   // 0 trace, 1 constraint, 2.. FRI

  // ---- batched hasher calls over (child, payload) lists
  // hash_elements of element strings of mixed lengths: one device batch per distinct length
  auto hash_groups = [&](const std::vector<std::pair<size_t, std::vector<zkl_fe>>> &items, std::vector<dig> &digests) -> int {
    digests.assign(items.size(), dig{});
    std::map<size_t, std::vector<size_t>> by_len;
    for (size_t k = 0; k < items.size(); k++) by_len[items[k].second.size()].push_back(k);
    for (auto &g : by_len) {
      std::vector<zkl_fe> flat; flat.reserve(g.first * g.second.size());
      for (size_t k : g.second) flat.insert(flat.end(), items[k].second.begin(), items[k].second.end());
      std::vector<uint8_t> o(32 * g.second.size());
      ZKL_TRY(zkl_hash_elements_batch(ctx, flat.data(), g.first, g.second.size(), o.data()));
      for (size_t j = 0; j < g.second.size(); j++) memcpy(digests[g.second[j]].data(), o.data() + 32 * j, 32);
    }
    return 0;
  };
  std::vector<dig> seed(count);
  auto reseed = [&](const std::vector<size_t> &ids, const std::vector<const uint8_t *> &digests) -> int {
    if (ids.empty()) return 0;
    std::vector<uint8_t> pairs(64 * ids.size()), o(32 * ids.size());
    for (size_t k = 0; k < ids.size(); k++) { memcpy(pairs.data() + 64 * k, seed[ids[k]].data(), 32); memcpy(pairs.data() + 64 * k + 32, digests[k], 32); }
    ZKL_TRY(zkl_merge_batch(ctx, pairs.data(), ids.size(), o.data()));
    for (size_t k = 0; k < ids.size(); k++) memcpy(seed[ids[k]].data(), o.data() + 32 * k, 32);
    return 0;
  };
  // merge_with_int(seed[id], value) for a list of (id, value)
  auto mwi = [&](const std::vector<std::pair<size_t, uint64_t>> &req, std::vector<dig> &res) -> int {
    res.assign(req.size(), dig{});
    if (req.empty()) return 0;
    std::vector<uint8_t> seeds(32 * req.size()), o(32 * req.size());
    std::vector<uint64_t> vals(req.size());
    for (size_t k = 0; k < req.size(); k++) { memcpy(seeds.data() + 32 * k, seed[req[k].first].data(), 32); vals[k] = req[k].second; }
    ZKL_TRY(zkl_merge_with_int_pairs(ctx, seeds.data(), vals.data(), req.size(), o.data()));
    for (size_t k = 0; k < req.size(); k++) memcpy(res[k].data(), o.data() + 32 * k, 32);
    return 0;
  };
  auto fe_of = [](const dig &d) { zkl_fe x; memcpy(&x.lo, d.data(), 8); memcpy(&x.hi, d.data() + 8, 8); return x; };
  auto fe_at = [](const uint8_t *p) { zkl_fe x; memcpy(&x.lo, p, 8); memcpy(&x.hi, p + 8, 8); return x; };
  auto mkfe = [](uint64_t lo) { zkl_fe x; x.lo = lo; x.hi = 0; return x; };
  auto live_ids = [&]() { std::vector<size_t> ids; for (size_t i = 0; i < count; i++) if (live(i)) ids.push_back(i); return ids; };

  // ---- 1. coin seeds: Context::to_elements() ++ AirPublicInputs::to_elements() (agg/fs.rs:69-73)
  {
    std::vector<std::pair<size_t, std::vector<zkl_fe>>> items;
    for (size_t i : live_ids()) {
      std::vector<zkl_fe> e;
      e.push_back(mkfe(((uint64_t)v[i].W << 8) | v[i].aux)); e.push_back(mkfe(v[i].n()));
      zkl_fe mlo; mlo.lo = 0xFFFFD30000000001ull; mlo.hi = 0; zkl_fe mhi; mhi.lo = 0xFFFFFFFFFFFFFFFFull; mhi.hi = 0;
      e.push_back(mlo); e.push_back(mhi);
      e.push_back(mkfe(v[i].num_constraints));
      e.push_back(mkfe((((uint64_t)v[i].ext << 8 | v[i].folding) << 8) | v[i].rem_deg));
      e.push_back(mkfe(v[i].grind)); e.push_back(mkfe(v[i].blowup)); e.push_back(mkfe(v[i].q));
      e.insert(e.end(), children[i].pi_elems, children[i].pi_elems + children[i].num_pi);
      items.push_back({i, std::move(e)});
    }
    std::vector<dig> d; ZKL_TRY(hash_groups(items, d));
    for (size_t k = 0; k < items.size(); k++) seed[items[k].first] = d[k];
  }
  // ---- 2. trace root, constraint root, OOD point
  for (size_t which = 0; which < 2; which++) {
    std::vector<size_t> ids = live_ids(); std::vector<const uint8_t *> d;
    for (size_t i : ids) d.push_back(root_of(i, which));
    ZKL_TRY(reseed(ids, d));
  }
  {
    std::vector<std::pair<size_t, uint64_t>> req; for (size_t i : live_ids()) req.push_back({i, 1});
    std::vector<dig> res; ZKL_TRY(mwi(req, res));
    for (size_t k = 0; k < req.size(); k++) out[req[k].first].z = fe_of(res[k]);
  }
  // ---- 3. OOD frame digest: current(z) || quotient(z) || next(zg) || quotient(zg)
  {
    std::vector<std::pair<size_t, std::vector<zkl_fe>>> items;
    for (size_t i : live_ids()) {
      const size_t W = v[i].W, C = v[i].ood_cons.n / 32;
      std::vector<zkl_fe> e;
      for (size_t c = 0; c < W; c++) e.push_back(fe_at(v[i].ood_trace.p + 16 * c));
      for (size_t c = 0; c < C; c++) e.push_back(fe_at(v[i].ood_cons.p + 16 * c));
      for (size_t c = 0; c < W; c++) e.push_back(fe_at(v[i].ood_trace.p + 16 * (W + c)));
      for (size_t c = 0; c < C; c++) e.push_back(fe_at(v[i].ood_cons.p + 16 * (C + c)));
      items.push_back({i, std::move(e)});
    }
    std::vector<dig> d; ZKL_TRY(hash_groups(items, d));
    std::vector<size_t> ids; std::vector<const uint8_t *> dp;
    for (size_t k = 0; k < items.size(); k++) { ids.push_back(items[k].first); dp.push_back(d[k].data()); }
    ZKL_TRY(reseed(ids, dp));
  }
  // ---- 4. DEEP coefficients: one per trace column, then one per composition column
  {
    std::vector<std::pair<size_t, uint64_t>> req;
    for (size_t i : live_ids()) { out[i].num_deep = (uint32_t)(v[i].W + v[i].ood_cons.n / 32); for (uint32_t c = 1; c <= out[i].num_deep; c++) req.push_back({i, c}); }
    std::vector<dig> res; ZKL_TRY(mwi(req, res));
    for (size_t k = 0; k < req.size(); k++) out_deep[req[k].first * deep_stride + (req[k].second - 1)] = fe_of(res[k]);
  }
  // ---- 5. FRI alphas: a transcript takes part in round d while it still has a root
  {
    size_t rounds = 0;
    for (size_t i : live_ids()) rounds = std::max(rounds, v[i].num_fri_layers() + 1);
    std::vector<uint64_t> degree(count);
    for (size_t i = 0; i < count; i++) degree[i] = live(i) ? v[i].n() : 0;
    for (size_t d = 0; d < rounds; d++) {
      std::vector<size_t> ids; std::vector<const uint8_t *> dp;
      for (size_t i : live_ids()) if (d < v[i].num_fri_layers() + 1) { ids.push_back(i); dp.push_back(root_of(i, 2 + d)); }
      ZKL_TRY(reseed(ids, dp));
      std::vector<std::pair<size_t, uint64_t>> req; for (size_t i : ids) req.push_back({i, 1});
      std::vector<dig> res; ZKL_TRY(mwi(req, res));
      for (size_t k = 0; k < ids.size(); k++) {
        const size_t i = ids[k];
        out[i].alphas[out[i].num_alphas++] = fe_of(res[k]);
        if (d != v[i].num_fri_layers() && degree[i] % v[i].folding) st[i] = RP_FRI_DEGREE;
        degree[i] /= v[i].folding;
      }
    }
  }
  // ---- 6. proof of work, query positions (sorted, de-duplicated)
  std::vector<std::vector<uint32_t>> pos(count);
  {
    std::vector<std::pair<size_t, uint64_t>> req; for (size_t i : live_ids()) req.push_back({i, v[i].nonce});
    std::vector<dig> res; ZKL_TRY(mwi(req, res));
    for (size_t k = 0; k < req.size(); k++) {
      const size_t i = req[k].first;
      uint64_t head; memcpy(&head, res[k].data(), 8);
      const uint32_t tz = head == 0 ? 64 : (uint32_t)__builtin_ctzll(head);
      if (tz < v[i].grind) st[i] = RP_POW; else seed[i] = res[k];      // draw_integers: seed' = merge_with_int(seed, nonce)
      out[i].pow_nonce = v[i].nonce;
    }
    req.clear();
    for (size_t i : live_ids()) for (uint32_t c = 1; c <= v[i].q; c++) req.push_back({i, c});
    ZKL_TRY(mwi(req, res));
    for (size_t k = 0; k < req.size(); k++) { uint64_t x; memcpy(&x, res[k].data(), 8); pos[req[k].first].push_back((uint32_t)(x & (v[req[k].first].N() - 1))); }
    for (size_t i : live_ids()) {
      std::sort(pos[i].begin(), pos[i].end()); pos[i].erase(std::unique(pos[i].begin(), pos[i].end()), pos[i].end());
      if (pos[i].size() != v[i].unique || pos[i].size() > 256) { st[i] = RP_UNIQUE; continue; }
      out[i].num_positions = (uint32_t)pos[i].size();
      for (size_t k = 0; k < pos[i].size(); k++) out[i].positions[k] = pos[i][k];
    }
  }

  // ---- 7. openings: leaves in batches, every tree in one merge plan
  if (verify_openings) {
    // row leaves of a (k x width) element matrix, hash_elements per partition and merge_many across partitions (agg/child.rs:1025-1045)
    struct row_job { size_t child; const uint8_t *values; size_t rows, width, psize; };
    auto row_leaves = [&](const std::vector<row_job> &jobs, std::vector<std::vector<dig>> &leaves) -> int {
      leaves.assign(jobs.size(), {});
      std::map<std::pair<size_t, size_t>, std::vector<size_t>> by_shape;
      for (size_t j = 0; j < jobs.size(); j++) by_shape[{jobs[j].width, jobs[j].psize}].push_back(j);
      for (auto &g : by_shape) {
        const size_t width = g.first.first, ps = g.first.second, nchunks = (width + ps - 1) / ps;
        size_t R = 0; for (size_t j : g.second) R += jobs[j].rows;
        std::vector<std::vector<uint8_t>> part(nchunks);
        for (size_t qi = 0; qi < nchunks; qi++) {
          const size_t first = qi * ps, len = std::min(ps, width - first);
          std::vector<zkl_fe> flat; flat.reserve(R * len);
          for (size_t j : g.second) for (size_t r = 0; r < jobs[j].rows; r++) for (size_t c = 0; c < len; c++) flat.push_back(fe_at(jobs[j].values + 16 * (r * width + first + c)));
          part[qi].resize(32 * R);
          ZKL_TRY(zkl_hash_elements_batch(ctx, flat.data(), len, R, part[qi].data()));
        }
        std::vector<uint8_t> leaf(32 * R);
        if (nchunks == 1) leaf = part[0];
        else {  // leaf = merge_many(partition digests): partition-major digest matrix on the device, one launch for all rows
          std::vector<fe> h(nchunks * R);
          for (size_t qi = 0; qi < nchunks; qi++) for (size_t r = 0; r < R; r++) h[qi * R + r] = fold_digest_host(part[qi].data() + 32 * r);
          dev_buf<fe> d, o; ZKL_CHECK(ctx, d.alloc(h.size())); ZKL_CHECK(ctx, o.alloc(R));
          ZKL_CHECK(ctx, cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
          ZKL_TRY(dev_merge_partitions(ctx, d.p, R, R, nchunks, o.p));
          ZKL_TRY(download_digests(ctx, o.p, R, leaf.data()));
        }
        size_t at = 0;
        for (size_t j : g.second) { leaves[j].resize(jobs[j].rows); for (size_t r = 0; r < jobs[j].rows; r++, at++) memcpy(leaves[j][r].data(), leaf.data() + 32 * at, 32); }
      }
      return 0;
    };
    std::vector<row_job> tjobs, cjobs;
    for (size_t i : live_ids()) {
      const size_t k = pos[i].size();
      if (k == 0 || v[i].tq_values.n != 16 * k * v[i].W || v[i].cq_values.n == 0 || v[i].cq_values.n % (16 * k)) { st[i] = RP_MALFORMED; continue; }
      const size_t C = v[i].cq_values.n / (16 * k);
      if (v[i].layers.size() != v[i].num_fri_layers() || v[i].remainder.n == 0 || v[i].remainder.n % 16) { st[i] = RP_MALFORMED; continue; }
      zkl_options o; memset(&o, 0, sizeof o); o.num_partitions = v[i].parts; o.hash_rate = v[i].rate;
      tjobs.push_back({i, v[i].tq_values.p, k, v[i].W, partition_size_of(&o, v[i].W)});
      cjobs.push_back({i, v[i].cq_values.p, k, C, partition_size_of(&o, C)});
    }
    std::vector<std::vector<dig>> tleaves, cleaves;
    ZKL_TRY(row_leaves(tjobs, tleaves)); ZKL_TRY(row_leaves(cjobs, cleaves));
    // FRI layer pairs of every child and layer in one batch; remainders grouped by length
    struct fri_job { size_t child, layer; std::vector<uint32_t> positions; size_t first_leaf; };
    std::vector<fri_job> fjobs; std::vector<zkl_fe> pairs;
    std::vector<std::pair<size_t, std::vector<zkl_fe>>> rems;
    for (size_t i : live_ids()) {
      std::vector<uint32_t> p = pos[i]; size_t domain = v[i].N();
      bool ok = true;
      const size_t first_job = fjobs.size(), first_pair = pairs.size();
      for (size_t d = 0; d < v[i].layers.size() && ok; d++) {
        std::vector<uint32_t> f;   // agg/child.rs:1072-1099: positions mod the folded domain, first occurrence order
        for (uint32_t x : p) { uint32_t y = (uint32_t)(x % (domain / 2)); if (std::find(f.begin(), f.end(), y) == f.end()) f.push_back(y); }
        if (v[i].layers[d].first.n != 32 * f.size()) { ok = false; break; }
        fjobs.push_back({i, d, f, pairs.size() / 2});
        for (size_t e = 0; e < 2 * f.size(); e++) pairs.push_back(fe_at(v[i].layers[d].first.p + 16 * e));
        p = f; domain /= 2;
      }
      if (!ok) { st[i] = RP_MALFORMED; fjobs.resize(first_job); pairs.resize(first_pair); continue; }
      std::vector<zkl_fe> r; for (size_t e = 0; e < v[i].remainder.n / 16; e++) r.push_back(fe_at(v[i].remainder.p + 16 * e));
      rems.push_back({i, std::move(r)});
    }
    std::vector<uint8_t> fleaf(32 * (pairs.size() / 2));
    if (!pairs.empty()) ZKL_TRY(zkl_hash_elements_batch(ctx, pairs.data(), 2, pairs.size() / 2, fleaf.data()));
    {
      std::vector<dig> d; ZKL_TRY(hash_groups(rems, d));
      for (size_t k = 0; k < rems.size(); k++) {
        const size_t i = rems[k].first;
        if (live(i) && memcmp(d[k].data(), root_of(i, 2 + v[i].num_fri_layers()), 32)) st[i] = RP_REMAINDER;
      }
    }
    rp_plan plan;
    auto add = [&](size_t i, rp_span paths, const std::vector<const uint8_t *> &leaves, const std::vector<uint32_t> &positions, const uint8_t *expected, int status) {
      int depth; std::vector<std::vector<const uint8_t *>> lists;
      if (!rp_read_batch_proof(paths, depth, lists) || !plan.add_opening(depth, lists, leaves, positions, expected, i, status)) { if (live(i)) st[i] = RP_BATCH_PROOF; }
    };
    for (size_t j = 0; j < tjobs.size(); j++) {
      const size_t i = tjobs[j].child;
      if (!live(i)) continue;
      std::vector<const uint8_t *> lt, lc;
      for (auto &d : tleaves[j]) lt.push_back(d.data());
      for (auto &d : cleaves[j]) lc.push_back(d.data());
      add(i, v[i].tq_paths, lt, pos[i], root_of(i, 0), RP_TRACE_OPENING);
      add(i, v[i].cq_paths, lc, pos[i], root_of(i, 1), RP_CONSTRAINT_OPENING);
    }
    for (auto &fj : fjobs) {
      if (!live(fj.child)) continue;
      std::vector<const uint8_t *> lf;
      for (size_t e = 0; e < fj.positions.size(); e++) lf.push_back(fleaf.data() + 32 * (fj.first_leaf + e));
      add(fj.child, v[fj.child].layers[fj.layer].second, lf, fj.positions, root_of(fj.child, 2 + fj.layer), RP_FRI_OPENING | (int)(fj.layer << 8));
    }
    if (!plan.roots.empty()) {
      std::vector<uint8_t> known(32 * plan.known.size());
      for (size_t k = 0; k < plan.known.size(); k++) memcpy(known.data() + 32 * k, plan.known[k], 32);
      std::vector<uint32_t> base(plan.levels.size()), counts, left, right, want;
      uint32_t acc = (uint32_t)plan.known.size();
      for (size_t l = 0; l < plan.levels.size(); l++) { base[l] = acc; acc += (uint32_t)plan.levels[l].size(); counts.push_back((uint32_t)plan.levels[l].size()); }
      auto index = [&](rp_node x) { return x.level < 0 ? x.idx : base[(size_t)x.level] + x.idx; };
      for (auto &lv : plan.levels) for (auto &m : lv) { left.push_back(index(m.first)); right.push_back(index(m.second)); }
      for (auto &r : plan.roots) want.push_back(index(r.node));
      std::vector<uint8_t> got(32 * want.size());
      ZKL_TRY(zkl_merge_plan(ctx, known.data(), plan.known.size(), left.data(), right.data(), counts.data(), counts.size(), want.data(), want.size(), got.data()));
      for (size_t k = 0; k < plan.roots.size(); k++) {
        auto &r = plan.roots[k];
        if (live(r.child) && memcmp(got.data() + 32 * k, r.expected, 32)) st[r.child] = r.status;
      }
    }
  }
  for (size_t i = 0; i < count; i++) out[i].status = st[i];
  return 0;
}
#endif  // ZKL_REPLAY_HOST_ONLY
