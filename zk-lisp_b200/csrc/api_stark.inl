// api_stark.inl — hook-level entry points and the whole-segment prover (included by zkl_b200.cu).

// ------------------------------------------------------------------------------------------------
// committed matrix = coset-major LDE + Merkle tree (DefaultTraceLde / DefaultConstraintCommitment)
// ------------------------------------------------------------------------------------------------
struct committed_matrix {
  size_t width = 0, n = 0, blowup = 0, N = 0;
  dev_buf<fe> polys;   // width x n coefficients
  dev_buf<fe> lde;     // width x N, coset-major
  dev_buf<fe> leaves;  // N digests (as elements), natural row order
  dev_buf<fe> nodes;   // N entries, nodes[1] = root
  dev_buf<fe> scratch;
  size_t ext = 1;      // rows are E elements stored as `ext` base columns each ((a, 0) pairs for the quadratic extension)
};
struct zkl_trace_lde { committed_matrix m; };
struct zkl_constraint_commitment { committed_matrix m; };

// polys must already hold width x n coefficients
// kev: optional 3 events recorded before the LDE kernel, between LDE and row hashing, and after row hashing
static int commit_matrix(zkl_ctx *ctx, committed_matrix &m, const zkl_options *o, cudaEvent_t *kev = nullptr) {
  ZKL_CHECK(ctx, m.lde.alloc(m.width * m.N));
  ZKL_CHECK(ctx, m.leaves.alloc(m.N));
  ZKL_CHECK(ctx, m.nodes.alloc(m.N));
  if (kev) ZKL_CHECK(ctx, cudaEventRecord(kev[0], ctx->stream));
  ZKL_TRY(dev_lde(ctx, m.polys.p, m.lde.p, m.width, m.n, m.blowup));
  if (kev) ZKL_CHECK(ctx, cudaEventRecord(kev[1], ctx->stream));
  ZKL_TRY(dev_hash_rows(ctx, m.lde.p, m.width, m.n, m.blowup, o, m.leaves.p, m.scratch, m.ext));
  if (kev) ZKL_CHECK(ctx, cudaEventRecord(kev[2], ctx->stream));
  ZKL_TRY(dev_merkle(ctx, m.leaves.p, m.nodes.p, m.N));
  return 0;
}
static int check_opts(zkl_ctx *ctx, const zkl_options *o, size_t n) {
  if (o->fri_folding != 2) { ctx->set_error("only FRI folding factor 2 is supported (prove.rs:968)"); return 1; }
  if (o->blowup < 2 || (o->blowup & (o->blowup - 1)) || o->blowup > 128) { ctx->set_error("blowup must be a power of two in [2,128]"); return 1; }
  if (o->field_extension > 2) { ctx->set_error("field_extension must be 0/1 (None) or 2 (Quadratic); zk-lisp never asks for Cubic (prove.rs:647-651,967)"); return 1; }
  if (n < 8 || n > NTT_MAX || (n & (n - 1))) { ctx->set_error("trace length must be a power of two in [8, 2^24]"); return 1; }
  // the FRI remainder is interpolated by one CTA (k_fri_remainder): its domain (max degree + 1) * blowup is bounded
  if ((size_t)(o->fri_remainder_max_degree + 1) * o->blowup > 256) { ctx->set_error("FRI remainder domain (fri_remainder_max_degree + 1) * blowup must be <= 256"); return 1; }
  return 0;
}

static int trace_commit_dev(zkl_ctx *ctx, const fe *trace_dev, size_t width, size_t n, const zkl_options *o, zkl_trace_lde *h,
                            cudaEvent_t *kev = nullptr) {
  committed_matrix &m = h->m;
  m.width = width; m.n = n; m.blowup = o->blowup; m.N = n * o->blowup;
  ZKL_CHECK(ctx, m.polys.alloc(width * n));
  ZKL_TRY(dev_interpolate(ctx, trace_dev, m.polys.p, width, n));
  return commit_matrix(ctx, m, o, kev);
}

extern "C" int zkl_trace_lde_commit(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n, const zkl_options *o,
                                    uint8_t out_root[32], zkl_fe *out_polys, zkl_trace_lde **out_handle) {
  ZKL_TRY(need_suite(ctx));
  ZKL_TRY(check_opts(ctx, o, n));
  zkl_trace_lde *h = new zkl_trace_lde();
  dev_buf<fe> tr;
  int rc = 0;
  do {
    if (tr.alloc(width * n) != cudaSuccess) { ctx->set_error("cudaMalloc(trace) failed"); rc = 1; break; }
    if (cudaMemcpyAsync(tr.p, trace, width * n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ctx->set_error("H2D failed"); rc = 1; break; }
    if ((rc = trace_commit_dev(ctx, tr.p, width, n, o, h))) break;
    fe root;
    if (cudaMemcpyAsync(&root, h->m.nodes.p + 1, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ctx->set_error("D2H failed"); rc = 1; break; }
    if (out_polys && cudaMemcpyAsync(out_polys, h->m.polys.p, width * n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ctx->set_error("D2H failed"); rc = 1; break; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->set_error(std::string("trace commit: ") + cudaGetErrorString(cudaGetLastError())); rc = 1; break; }
    digest_from_fe_host(root, out_root);
  } while (0);
  if (rc) { delete h; return rc; }
  *out_handle = h;
  return 0;
}
extern "C" void zkl_trace_lde_free(zkl_ctx *ctx, zkl_trace_lde *h) {
  if (!h) return;
  enter(ctx); cudaStreamSynchronize(ctx->stream);
  delete h;
}
extern "C" int zkl_trace_lde_read_frame(zkl_ctx *ctx, const zkl_trace_lde *h, size_t lde_step, zkl_fe *out_cur, zkl_fe *out_next) {
  ZKL_TRY(enter(ctx));
  const committed_matrix &m = h->m;
  if (lde_step >= m.N) { ctx->set_error("lde_step out of range"); return 1; }
  uint32_t pos[2] = {(uint32_t)lde_step, (uint32_t)((lde_step + m.blowup) % m.N)};
  dev_buf<uint32_t> dp; dev_buf<fe> rows; ZKL_CHECK(ctx, dp.alloc(2)); ZKL_CHECK(ctx, rows.alloc(2 * m.width));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dp.p, pos, sizeof pos, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_gather_rows, ceil_div(2 * m.width, 128), 128, 0, m.lde.p, m.N, m.n, (int)m.blowup, (int)m.width, dp.p, 2, rows.p);
  ZKL_CHECK(ctx, cudaMemcpyAsync(out_cur, rows.p, m.width * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(out_next, rows.p + m.width, m.width * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Merkle batch openings (winter-crypto MerkleTree::prove_batch [WF-mem]) — pure index logic on the host.
// Returns, per normalized leaf pair, the list of node references to ship: (is_leaf, index).
// ------------------------------------------------------------------------------------------------
struct node_ref { bool leaf; uint32_t index; };
static std::vector<std::vector<node_ref>> batch_proof_plan(const std::vector<uint32_t> &indexes, size_t n_leaves) {
  int depth = ilog2_sz(n_leaves);
  std::vector<uint32_t> norm;
  for (uint32_t i : indexes) norm.push_back(i & ~1u);
  std::sort(norm.begin(), norm.end());
  norm.erase(std::unique(norm.begin(), norm.end()), norm.end());
  std::vector<std::vector<node_ref>> lists(norm.size());
  std::vector<size_t> next;
  for (size_t i = 0; i < norm.size(); i++) {
    for (uint32_t j = norm[i]; j < norm[i] + 2; j++)
      if (std::find(indexes.begin(), indexes.end(), j) == indexes.end()) lists[i].push_back({true, j});
    next.push_back(((size_t)norm[i] + n_leaves) >> 1);
  }
  for (int d = 1; d < depth; d++) {
    std::vector<size_t> cur = next; next.clear();
    size_t i = 0;
    while (i < cur.size()) {
      size_t sib = cur[i] ^ 1;
      if (i + 1 < cur.size() && cur[i + 1] == sib) i++;
      else lists[i].push_back({false, (uint32_t)sib});
      next.push_back(sib >> 1);
      i++;
    }
  }
  return lists;
}
// fetch the planned nodes from device (leaves / nodes arrays of elements) into digests
static int fetch_batch_proof(zkl_ctx *ctx, const fe *leaves_dev, const fe *nodes_dev, const std::vector<std::vector<node_ref>> &plan,
                             std::vector<std::vector<fe>> &out) {
  // single gather over a concatenated index list: leaves and nodes live in separate buffers -> two gathers
  std::vector<uint32_t> li, ni;
  for (auto &l : plan) for (auto &r : l) (r.leaf ? li : ni).push_back(r.index);
  std::vector<fe> lv(li.size()), nv(ni.size());
  dev_buf<uint32_t> di; dev_buf<fe> dv;
  auto gather = [&](const fe *src, std::vector<uint32_t> &idx, std::vector<fe> &dst) -> int {
    if (idx.empty()) return 0;
    ZKL_CHECK(ctx, di.alloc(idx.size())); ZKL_CHECK(ctx, dv.alloc(idx.size()));
    ZKL_CHECK(ctx, cudaMemcpyAsync(di.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather_fe, ceil_div(idx.size(), 128), 128, 0, src, di.p, (int)idx.size(), dv.p);
    ZKL_CHECK(ctx, cudaMemcpyAsync(dst.data(), dv.p, idx.size() * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
  };
  ZKL_TRY(gather(leaves_dev, li, lv));
  ZKL_TRY(gather(nodes_dev, ni, nv));
  size_t a = 0, b = 0;
  out.clear();
  for (auto &l : plan) {
    std::vector<fe> v;
    for (auto &r : l) v.push_back(r.leaf ? lv[a++] : nv[b++]);
    out.push_back(v);
  }
  return 0;
}

static int query_matrix(zkl_ctx *ctx, const committed_matrix &m, const uint32_t *positions, size_t k, zkl_fe *out_rows,
                        zkl_batch_proof *bp) {
  ZKL_TRY(enter(ctx));
  std::vector<uint32_t> pos(positions, positions + k);
  for (auto p : pos) if (p >= m.N) { ctx->set_error("query position out of range"); return 1; }
  dev_buf<uint32_t> dp; dev_buf<fe> rows; ZKL_CHECK(ctx, dp.alloc(k)); ZKL_CHECK(ctx, rows.alloc(k * m.width));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dp.p, pos.data(), k * 4, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_gather_rows, ceil_div(k * m.width, 128), 128, 0, m.lde.p, m.N, m.n, (int)m.blowup, (int)m.width, dp.p, (int)k, rows.p);
  ZKL_CHECK(ctx, cudaMemcpyAsync(out_rows, rows.p, k * m.width * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  auto plan = batch_proof_plan(pos, m.N);
  std::vector<std::vector<fe>> lists;
  ZKL_TRY(fetch_batch_proof(ctx, m.leaves.p, m.nodes.p, plan, lists));
  bp->depth = (uint8_t)ilog2_sz(m.N);
  bp->num_lists = (uint32_t)lists.size();
  uint32_t total = 0;
  for (size_t i = 0; i < lists.size(); i++) {
    bp->list_len[i] = (uint32_t)lists[i].size();
    for (auto &d : lists[i]) { digest_from_fe_host(d, bp->nodes + 32 * (size_t)total); total++; }
  }
  bp->total_nodes = total;
  return 0;
}
extern "C" int zkl_trace_query(zkl_ctx *ctx, const zkl_trace_lde *h, const uint32_t *positions, size_t k, zkl_fe *out_rows,
                               zkl_batch_proof *out_proof) {
  return query_matrix(ctx, h->m, positions, k, out_rows, out_proof);
}
extern "C" int zkl_constraint_query(zkl_ctx *ctx, const zkl_constraint_commitment *h, const uint32_t *positions, size_t k,
                                    zkl_fe *out_rows, zkl_batch_proof *out_proof) {
  return query_matrix(ctx, h->m, positions, k, out_rows, out_proof);
}

// ------------------------------------------------------------------------------------------------
// constraint evaluation (hook 2)
// ------------------------------------------------------------------------------------------------
// AIR parameters on the device. Nothing is cached between calls: the constants (11 KB) and the assertion table (32 B per
// assertion, 0.58 MB for a 4096-row segment) are rebuilt in the context's pinned staging buffer and uploaded per call, so a
// segment can never be proven with another instance's constants. Assertions are ordered as Winterfell orders them (by step,
// then column: single assertions have stride 0 [WF-mem]) with two counting sorts.
struct air_device_state {
  dev_buf<uint8_t> blob;   // air_dev | assertion_dev[na] | group_start[ng + 1] | slot_col[nslots]
  struct view { air_dev *p; } air;
  struct view_as { assertion_dev *p; } as;
  struct view_u32 { uint32_t *p; } group_start, slot_col;
  dev_buf<fe> coset_consts, u, bpoly;
  dev_buf<uint8_t> tmpl_blob;   // closed-form assertion list (host_air.inl: sched_assertions) the device expands
  int nslots = 0, ngroups = 0;
};
// Host wait for everything enqueued on the context's stream (ZKL_WAIT_MODE / ctx->wait_mode): 1 = spin (cudaStreamSynchronize:
// lowest latency, one core per waiting prover), 2 = block on a cudaEventBlockingSync event (no core, ~0.3 ms of wake-up latency per
// wait), 3 = poll the event and sched_yield() between polls. 0 = automatic: spin while this process has at least two cores per
// proving context, poll-and-yield otherwise (a pool of provers or several ranks sharing the host's cores).
static int ctx_wait(zkl_ctx *ctx) {
  // automatic: block when this process runs more proving contexts than half the cores it may use (spinning waiters would then
  // take the cores the other provers' host code needs); a blocking wait costs ~0.3 ms of wake-up latency per proof otherwise
  const bool crowded = 2 * g_ctx_total.load(std::memory_order_relaxed) > g_host_cores;
  const int mode = ctx->wait_mode ? ctx->wait_mode : (crowded ? 3 : 1);
  if (mode == 1) { ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }
  if (!ctx->wait_ev) ZKL_CHECK(ctx, cudaEventCreateWithFlags(&ctx->wait_ev, cudaEventBlockingSync | cudaEventDisableTiming));
  ZKL_CHECK(ctx, cudaEventRecord(ctx->wait_ev, ctx->stream));
  if (mode == 2) { ZKL_CHECK(ctx, cudaEventSynchronize(ctx->wait_ev)); return 0; }
  // mode 3: poll and yield — no wake-up latency while cores are free, and a waiting prover hands its core to any runnable one
  for (;;) {
    cudaError_t e = cudaEventQuery(ctx->wait_ev);
    if (e == cudaSuccess) return 0;
    if (e != cudaErrorNotReady) { ZKL_CHECK(ctx, e); }
    sched_yield();
  }
}
// Pinned host staging, one region per context, handed out by a bump pointer that restarts with every entry point:
// parameter uploads and result downloads of a call never pass through pageable memory (a cudaMemcpyAsync on pageable memory
// is staged by the driver and may wait for the stream). A region may only be (re)used when nothing enqueued earlier still
// refers to it: every entry point that stages ends with a host wait.
static void *stage_alloc(zkl_ctx *ctx, size_t bytes) {
  const size_t off = (ctx->stage_used + 63) & ~(size_t)63;
  if (off + bytes > ctx->stage_cap) return nullptr;
  ctx->stage_used = off + bytes;
  return (uint8_t *)ctx->stage_h + off;
}
static int stage_reserve(zkl_ctx *ctx, size_t bytes) {
  ctx->stage_used = 0;
  if (ctx->stage_cap >= bytes) return 0;
  if (ctx->stage_h) { ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->stage_h); ctx->stage_h = nullptr; ctx->stage_cap = 0; }
  size_t cap = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
  ZKL_CHECK(ctx, cudaHostAlloc(&ctx->stage_h, cap, cudaHostAllocDefault));
  ctx->stage_cap = cap;
  return 0;
}
// ---- the assertion list from its closed form (host_air.inl: sched_assertions), in Winterfell's order (step, then column) ----
// Template entry j of level l sits at rank l T + j + (extras ordered before it); a template step's group starts at
// l T + first_j + (extras at earlier steps). The <= 32 extras and the groups of steps only extras name are ranked on the host.
struct tmpl_dev { uint32_t slot, pos, col, pad; fe value; };       // sorted by (pos, col)
struct tgroup_dev { uint32_t pos, first_j; };                        // distinct positions of the template
struct extra_fix { uint32_t rank, pad[3]; assertion_dev a; };
struct group_fix { uint32_t gidx, gstart; };
struct expand_args {
  uint32_t levels, period, T, G, E, XO, NGF, nt;
  const tmpl_dev *tmpl; const tgroup_dev *tg; const uint32_t *ex_step, *ex_col, *xo_step; const extra_fix *ef; const group_fix *gf;
};
__global__ void k_expand_assertions(expand_args a, assertion_dev *__restrict__ as, uint32_t *__restrict__ gs) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nT = (size_t)a.levels * a.T, nG = (size_t)a.levels * a.G;
  if (t < nT) {
    const uint32_t l = (uint32_t)(t / a.T), j = (uint32_t)(t - (size_t)l * a.T);
    const tmpl_dev x = a.tmpl[j];
    const uint32_t step = l * a.period + x.pos;
    uint32_t eb = 0;
    for (uint32_t e = 0; e < a.E; e++) eb += (a.ex_step[e] < step || (a.ex_step[e] == step && a.ex_col[e] < x.col)) ? 1 : 0;
    const size_t rank = t + eb;
    assertion_dev o; o.slot = x.slot; o.step = step; o.cc_index = a.nt + (uint32_t)rank; o.pad = 0; o.value = x.value;
    as[rank] = o;
    return;
  }
  t -= nT;
  if (t < nG) {
    const uint32_t l = (uint32_t)(t / a.G), k = (uint32_t)(t - (size_t)l * a.G);
    const uint32_t step = l * a.period + a.tg[k].pos;
    uint32_t eb = 0, xo = 0;
    for (uint32_t e = 0; e < a.E; e++) eb += a.ex_step[e] < step ? 1 : 0;
    for (uint32_t e = 0; e < a.XO; e++) xo += a.xo_step[e] < step ? 1 : 0;
    gs[t + xo] = l * a.T + a.tg[k].first_j + eb;
    return;
  }
  t -= nG;
  if (t < a.E) { as[a.ef[t].rank] = a.ef[t].a; return; }
  t -= a.E;
  if (t < a.NGF) gs[a.gf[t].gidx] = a.gf[t].gstart;
}
static int fill_air_header(zkl_ctx *ctx, const zkl_air_params *a, air_dev &h) {
  memset(&h, 0, sizeof h);
  h.cols = a->cols;
  h.f_poseidon = a->f_poseidon; h.f_vm = a->f_vm; h.f_sponge = a->f_sponge; h.f_merkle = a->f_merkle; h.f_ram = a->f_ram;
  h.rom_enabled = a->rom_enabled; h.vm_usage_mask = a->vm_usage_mask; h.ram_delta_clk_bits = a->ram_delta_clk_bits;
  h.num_transition = a->num_transition; h.air_kind = a->air_kind;
  if (a->air_kind > ZKL_AIR_AGG) { ctx->set_error("unknown air_kind"); return 1; }
  if (a->air_kind == ZKL_AIR_AGG && (a->cols.width != ZKL_AGG_WIDTH || a->num_transition != ZKL_AGG_TRANSITIONS)) {
    ctx->set_error("ZlAggAir has 31 columns and 24 transition constraints (agg/layout.rs, agg/air.rs:69-94)"); return 1;
  }
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) h.pose_mds[i][j] = fe_from_abi(a->pose.mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 12; j++) h.pose_rc[r][j] = fe_from_abi(a->pose.rc[r][j]);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) h.rom_mds[i][j] = fe_from_abi(a->rom.mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 3; j++) h.rom_rc[r][j] = fe_from_abi(a->rom.rc[r][j]);
  for (int k = 0; k < 59; k++) { h.rom_w0[k] = fe_from_abi(a->rom_w_enc0[k]); h.rom_w1[k] = fe_from_abi(a->rom_w_enc1[k]); }
  h.program_fe0 = fe_from_abi(a->program_fe0); h.merkle_root_fe = fe_from_abi(a->merkle_root_fe);
  return 0;
}
// the closed form: a few KB cross PCIe, the device writes the ordered list (k_expand_assertions)
static int upload_air_sched(zkl_ctx *ctx, const zkl_air_params *a, size_t n, air_device_state &st, const sched_assertions &sa) {
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t T = sa.tmpl.size(), E = sa.extras.size(), na = sa.count();
  if (na != a->num_assertions || (size_t)sa.levels * sa.period != n || !T || E > 64) { ctx->set_error("inconsistent closed-form assertion list"); return 1; }
  std::vector<zkl_assertion> tm = sa.tmpl, ex = sa.extras;
  auto by_step_col = [](const zkl_assertion &x, const zkl_assertion &y) { return x.step != y.step ? x.step < y.step : x.column < y.column; };
  std::sort(tm.begin(), tm.end(), by_step_col); std::sort(ex.begin(), ex.end(), by_step_col);
  for (const auto &x : tm) if (x.column >= a->cols.width || x.step >= sa.period) { ctx->set_error("assertion outside the trace"); return 1; }
  for (const auto &x : ex) if (x.column >= a->cols.width || x.step >= n) { ctx->set_error("assertion outside the trace"); return 1; }
  // distinct positions of the template, slots (dense numbering of the asserted columns)
  std::vector<tgroup_dev> tg;
  for (size_t j = 0; j < T; j++) if (!j || tm[j].step != tm[j - 1].step) tg.push_back(tgroup_dev{tm[j].step, (uint32_t)j});
  const size_t G = tg.size();
  int slot_of[256]; for (int c = 0; c < 256; c++) slot_of[c] = -1;
  uint32_t slot_col_h[256]; int nslots = 0;
  auto slot = [&](uint32_t col) { if (slot_of[col] < 0) { slot_of[col] = nslots; slot_col_h[nslots++] = col; } return (uint32_t)slot_of[col]; };
  // steps only extras name, the groups they open, the final sentinel
  auto tmpl_has_pos = [&](uint32_t pos) { for (auto &g : tg) if (g.pos == pos) return true; return false; };
  std::vector<uint32_t> xo;
  for (size_t e = 0; e < E; e++) if (!tmpl_has_pos(ex[e].step % sa.period) && (xo.empty() || xo.back() != ex[e].step)) xo.push_back(ex[e].step);
  const size_t XO = xo.size(), ng = (size_t)sa.levels * G + XO, NGF = XO + 1;
  const size_t o_t = al(sizeof(air_dev)), o_g = o_t + al(T * sizeof(tmpl_dev)), o_es = o_g + al(G * sizeof(tgroup_dev)), o_ec = o_es + al((E + 1) * 4),
               o_xo = o_ec + al((E + 1) * 4), o_ef = o_xo + al((XO + 1) * 4), o_gf = o_ef + al((E + 1) * sizeof(extra_fix)), o_sc = o_gf + al(NGF * sizeof(group_fix)),
               small = o_sc + al(256 * 4);
  uint8_t *hb = (uint8_t *)stage_alloc(ctx, small);
  if (!hb) { ZKL_TRY(stage_reserve(ctx, small)); hb = (uint8_t *)stage_alloc(ctx, small); }
  ZKL_TRY(fill_air_header(ctx, a, *(air_dev *)hb));
  tmpl_dev *td = (tmpl_dev *)(hb + o_t); tgroup_dev *tgd = (tgroup_dev *)(hb + o_g);
  uint32_t *es = (uint32_t *)(hb + o_es), *ec = (uint32_t *)(hb + o_ec), *xod = (uint32_t *)(hb + o_xo);
  extra_fix *ef = (extra_fix *)(hb + o_ef); group_fix *gf = (group_fix *)(hb + o_gf);
  for (size_t j = 0; j < T; j++) { td[j].slot = slot(tm[j].column); td[j].pos = tm[j].step; td[j].col = tm[j].column; td[j].pad = 0; td[j].value = fe_from_abi(tm[j].value); }
  memcpy(tgd, tg.data(), G * sizeof(tgroup_dev));
  for (size_t e = 0; e < E; e++) {
    const uint32_t lvl = ex[e].step / sa.period, pos = ex[e].step % sa.period;
    size_t before = 0;   // template entries of the extra's own level ordered before it
    for (size_t j = 0; j < T; j++) before += (tm[j].step < pos || (tm[j].step == pos && tm[j].column < ex[e].column)) ? 1 : 0;
    const size_t rank = (size_t)lvl * T + before + e;
    es[e] = ex[e].step; ec[e] = ex[e].column;
    memset(&ef[e], 0, sizeof ef[e]);
    ef[e].rank = (uint32_t)rank;
    ef[e].a.slot = slot(ex[e].column); ef[e].a.step = ex[e].step; ef[e].a.cc_index = a->num_transition + (uint32_t)rank; ef[e].a.value = fe_from_abi(ex[e].value);
  }
  for (size_t x = 0; x < XO; x++) {
    const uint32_t lvl = xo[x] / sa.period, pos = xo[x] % sa.period;
    size_t tsteps = 0, tents = 0, eb = 0;
    for (auto &g : tg) tsteps += g.pos < pos ? 1 : 0;
    for (size_t j = 0; j < T; j++) tents += tm[j].step < pos ? 1 : 0;
    for (size_t e = 0; e < E; e++) eb += ex[e].step < xo[x] ? 1 : 0;
    xod[x] = xo[x];
    gf[x].gidx = (uint32_t)((size_t)lvl * G + tsteps + x); gf[x].gstart = (uint32_t)((size_t)lvl * T + tents + eb);
  }
  gf[XO].gidx = (uint32_t)ng; gf[XO].gstart = (uint32_t)na;
  memcpy(hb + o_sc, slot_col_h, 256 * 4);
  // device: the small tables, then the expanded list in the layout the kernels read
  const size_t off_as = al(sizeof(air_dev)), off_gs = off_as + al(na * sizeof(assertion_dev)), off_sc = off_gs + al((ng + 1) * 4), total = off_sc + al(256 * 4);
  ZKL_CHECK(ctx, st.tmpl_blob.alloc(small)); ZKL_CHECK(ctx, st.blob.alloc(total));
  ZKL_CHECK(ctx, cudaMemcpyAsync(st.tmpl_blob.p, hb, small, cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(st.blob.p, st.tmpl_blob.p, sizeof(air_dev), cudaMemcpyDeviceToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(st.blob.p + off_sc, st.tmpl_blob.p + o_sc, 256 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  st.air.p = (air_dev *)st.blob.p; st.as.p = (assertion_dev *)(st.blob.p + off_as);
  st.group_start.p = (uint32_t *)(st.blob.p + off_gs); st.slot_col.p = (uint32_t *)(st.blob.p + off_sc);
  st.nslots = nslots; st.ngroups = (int)ng;
  const uint8_t *d = st.tmpl_blob.p;
  expand_args ea{sa.levels, sa.period, (uint32_t)T, (uint32_t)G, (uint32_t)E, (uint32_t)XO, (uint32_t)NGF, a->num_transition,
                 (const tmpl_dev *)(d + o_t), (const tgroup_dev *)(d + o_g), (const uint32_t *)(d + o_es), (const uint32_t *)(d + o_ec), (const uint32_t *)(d + o_xo),
                 (const extra_fix *)(d + o_ef), (const group_fix *)(d + o_gf)};
  const size_t threads = (size_t)sa.levels * (T + G) + E + NGF;
  LAUNCH(ctx, k_expand_assertions, ceil_div(threads, 256), 256, 0, ea, st.as.p, st.group_start.p);
  return 0;
}
static int upload_air(zkl_ctx *ctx, const zkl_air_params *a, size_t n, air_device_state &st, const sched_assertions *sched = nullptr) {
  if (a->cols.width > 255) { ctx->set_error("trace width must fit in u8 (winterfell TraceInfo)"); return 1; }
  if (sched) return upload_air_sched(ctx, a, n, st, *sched);
  if (!a->assertions && a->num_assertions) { ctx->set_error("zkl_air_params.assertions is NULL"); return 1; }
  const size_t na = a->num_assertions;
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t off_as = al(sizeof(air_dev)), off_gs = off_as + al(std::max<size_t>(na, 1) * sizeof(assertion_dev));
  // group count is known only after the pass over the steps: reserve the worst case (one group per assertion)
  const size_t off_sc = off_gs + al((na + 1) * 4), total = off_sc + al(256 * 4);
  uint8_t *hb = (uint8_t *)stage_alloc(ctx, total);
  if (!hb) { ZKL_TRY(stage_reserve(ctx, total)); hb = (uint8_t *)stage_alloc(ctx, total); }   // hook-level call: the region is free (each ends with a wait)
  air_dev &h = *(air_dev *)hb;
  ZKL_TRY(fill_air_header(ctx, a, h));
  // order[] = assertion indices sorted by (step, column), stable: counting sort by column, then by step
  for (size_t i = 0; i < na; i++)
    if (a->assertions[i].column >= a->cols.width || a->assertions[i].step >= n) { ctx->set_error("assertion outside the trace"); return 1; }
  std::vector<uint32_t> &by_col = ctx->sort_a, &order = ctx->sort_b, &cnt = ctx->sort_cnt;
  by_col.resize(na); order.resize(na);
  cnt.assign(257, 0);
  for (size_t i = 0; i < na; i++) cnt[a->assertions[i].column + 1]++;
  for (int c = 0; c < 256; c++) cnt[c + 1] += cnt[c];
  for (size_t i = 0; i < na; i++) by_col[cnt[a->assertions[i].column]++] = (uint32_t)i;
  cnt.assign(n + 1, 0);
  for (size_t i = 0; i < na; i++) cnt[a->assertions[i].step + 1]++;
  for (size_t r = 0; r < n; r++) cnt[r + 1] += cnt[r];
  for (size_t i = 0; i < na; i++) order[cnt[a->assertions[by_col[i]].step]++] = by_col[i];
  assertion_dev *as = (assertion_dev *)(hb + off_as);
  uint32_t *gs = (uint32_t *)(hb + off_gs), *slot_col = (uint32_t *)(hb + off_sc);
  int slot_of[256]; for (int c = 0; c < 256; c++) slot_of[c] = -1;
  int nslots = 0, ng = 0;
  for (size_t q = 0; q < na; q++) {
    const zkl_assertion &x = a->assertions[order[q]];
    if (slot_of[x.column] < 0) { slot_of[x.column] = nslots; slot_col[nslots++] = x.column; }
    as[q].slot = (uint32_t)slot_of[x.column]; as[q].step = x.step; as[q].cc_index = a->num_transition + (uint32_t)q; as[q].pad = 0;
    as[q].value = fe_from_abi(x.value);
    if (q == 0 || x.step != a->assertions[order[q - 1]].step) gs[ng++] = (uint32_t)q;
  }
  gs[ng] = (uint32_t)na;
  st.nslots = nslots; st.ngroups = ng;
  ZKL_CHECK(ctx, st.blob.alloc(total));
  ZKL_CHECK(ctx, cudaMemcpyAsync(st.blob.p, hb, total, cudaMemcpyHostToDevice, ctx->stream));
  st.air.p = (air_dev *)st.blob.p; st.as.p = (assertion_dev *)(st.blob.p + off_as);
  st.group_start.p = (uint32_t *)(st.blob.p + off_gs); st.slot_col.p = (uint32_t *)(st.blob.p + off_sc);
  return 0;
}

// cc_dev: num_transition + num_assertions coefficients on the device. ce_out: ce*n evaluations, coset-major.
// `rows` (column-sharded prover): the full-width rows of CE cosets kp0 .. kp0 + ncos - 1 as [W][ncos][n]; only those cosets are written.
struct ce_rows { const fe *rows; int kp0, ncos; };
static int dev_eval_constraints(zkl_ctx *ctx, const committed_matrix &t, const zkl_air_params *a, air_device_state &st, const fe *cc_dev,
                                fe *ce_out, const ce_rows *rows = nullptr) {
  size_t n = t.n, ce = a->ce_blowup, M = n * ce, N = t.N;
  if (ce > t.blowup || ce > 16 || (ce & (ce - 1))) { ctx->set_error("ce_blowup must be a power of two <= min(blowup, 16)"); return 1; }
  const bool agg = a->air_kind == ZKL_AIR_AGG;
  if (n < 32 && !agg) { ctx->set_error("trace length must be >= 32 (one schedule level)"); return 1; }
  domain_tables *tN, *tM, *tn;
  ZKL_TRY(get_tables(ctx, N, &tN)); ZKL_TRY(get_tables(ctx, M, &tM)); ZKL_TRY(get_tables(ctx, n, &tn));
  // periodic table: data independent, cached per (n, ce) (the 31 cycle-32 schedule columns of ZkLispAir; ZlAggAir has none)
  fe g = fe_from_u64(3);
  fe *per_tab = nullptr;
  if (!agg) {
    size_t key = (n << 8) | ce;
    auto it = ctx->periodic.find(key);
    if (it == ctx->periodic.end()) {
      fe *p; ZKL_CHECK(ctx, cudaMalloc((void **)&p, ce * 32 * 31 * sizeof(fe)));
      LAUNCH(ctx, k_periodic_table, ceil_div(ce * 32, 64), 64, 0, p, (int)ce, (int)n, tM->wpow, M, fe_pow(g, n / 32), fe_inv(fe_from_u64(32)));
      it = ctx->periodic.emplace(key, p).first;
    }
    per_tab = it->second;
  }
  // per-coset constants: x^n = g^n * w_ce^k' (data independent, cached per (n, ce))
  fe *coset_consts;
  {
    size_t key = (n << 8) | ce | ((size_t)1 << 60);
    auto it = ctx->periodic.find(key);
    if (it == ctx->periodic.end()) {
      std::vector<fe> cconst(2 * ce + 2);
      fe gn = fe_pow(g, n), wce = host_root_of_unity(ilog2_sz(ce)), wk = fe_one();
      for (size_t k = 0; k < ce; k++) {
        fe xn1 = fe_sub(fe_mul(gn, wk), fe_one());
        cconst[2 * k] = fe_inv(xn1); cconst[2 * k + 1] = xn1;
        wk = fe_mul(wk, wce);
      }
      cconst[2 * ce] = fe_inv(fe_from_u64(n)); cconst[2 * ce + 1] = fe_from_u64(n);
      fe *p; ZKL_CHECK(ctx, cudaMalloc((void **)&p, cconst.size() * sizeof(fe)));
      ZKL_CHECK(ctx, cudaMemcpyAsync(p, cconst.data(), cconst.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
      ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      it = ctx->periodic.emplace(key, p).first;
    }
    coset_consts = it->second;
  }
  fe *plast;
  {
    size_t key = (n << 8) | ce | ((size_t)1 << 59) | ((size_t)t.blowup << 40);
    auto it = ctx->periodic.find(key);
    if (it == ctx->periodic.end()) {
      fe *p; ZKL_CHECK(ctx, cudaMalloc((void **)&p, M * sizeof(fe)));
      LAUNCH(ctx, k_plast_table, ceil_div(M, 128), 128, 0, tN->wpow, N, (int)n, (int)t.blowup, (int)ce, coset_consts, p);
      it = ctx->periodic.emplace(key, p).first;
    }
    plast = it->second;
  }
  // boundary polynomials: scatter -> iNTT -> coset LDE over the CE domain
  size_t nb = (size_t)st.nslots + 1;
  const size_t bcos = rows ? (size_t)rows->ncos : ce;   // the column-sharded prover extends the boundary polynomials over ITS cosets only
  ZKL_CHECK(ctx, st.u.alloc(nb * n)); ZKL_CHECK(ctx, st.bpoly.alloc(nb * bcos * n));
  ZKL_CHECK(ctx, cudaMemsetAsync(st.u.p, 0, nb * n * sizeof(fe), ctx->stream));
  if (st.ngroups) LAUNCH(ctx, k_boundary_scatter, ceil_div(st.ngroups, 128), 128, 0, st.as.p, st.group_start.p, st.ngroups, cc_dev, tn->wpow, (int)n, st.nslots, st.u.p);
  ZKL_TRY(dev_interpolate(ctx, st.u.p, st.u.p, nb, n));
  if (!rows) ZKL_TRY(dev_lde(ctx, st.u.p, st.bpoly.p, nb, n, ce));
  else ZKL_TRY(dev_lde_cosets(ctx, st.u.p, st.bpoly.p, nb, n, ce, (size_t)rows->kp0, bcos));
  boundary_dev bd{st.slot_col.p, st.bpoly.p, st.nslots, bcos * n, rows ? 1 : 0};
  if (!rows) LAUNCH(ctx, k_eval_constraints, ceil_div(M, 128), 128, 0, st.air.p, t.lde.p, N, (int)n, (int)t.blowup, (int)ce, per_tab, tN->wpow, cc_dev, bd,
                    coset_consts, plast, ce_out, N, 0, (int)ce, (int)(t.blowup / ce));
  else LAUNCH(ctx, k_eval_constraints, ceil_div((size_t)rows->ncos * n, 128), 128, 0, st.air.p, rows->rows, N, (int)n, (int)t.blowup, (int)ce, per_tab, tN->wpow,
              cc_dev, bd, coset_consts, plast, ce_out, (size_t)rows->ncos * n, rows->kp0, rows->ncos, 1);
  return 0;
}

// coset-major CE evaluations -> natural order (CompositionPolyTrace order) and back
__global__ void k_ce_to_natural(const fe *__restrict__ cm, fe *__restrict__ nat, int n, int ce, int to_natural) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ce) return;
  int kp = t / n, i = t - kp * n;
  int j = kp + ce * i;
  if (to_natural) st_fe(nat + j, ld_fe(cm + t)); else st_fe(nat + t, ld_fe(cm + j));
}

extern "C" int zkl_eval_constraints(zkl_ctx *ctx, const zkl_trace_lde *h, const zkl_air_params *air, const zkl_fe *coeffs, zkl_fe *out_ce) {
  ZKL_TRY(need_suite(ctx));
  const committed_matrix &t = h->m;
  if (air->cols.width != t.width) { ctx->set_error("air width does not match the committed trace"); return 1; }
  air_device_state st;
  ZKL_TRY(upload_air(ctx, air, t.n, st));
  size_t ncc = (size_t)air->num_transition + air->num_assertions, M = t.n * air->ce_blowup;
  dev_buf<fe> cc, ce, nat; ZKL_CHECK(ctx, cc.alloc(ncc)); ZKL_CHECK(ctx, ce.alloc(M)); ZKL_CHECK(ctx, nat.alloc(M));
  ZKL_CHECK(ctx, cudaMemcpyAsync(cc.p, coeffs, ncc * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  ZKL_TRY(dev_eval_constraints(ctx, t, air, st, cc.p, ce.p));
  LAUNCH(ctx, k_ce_to_natural, ceil_div(M, 256), 256, 0, ce.p, nat.p, (int)t.n, (int)air->ce_blowup, 1);
  ZKL_CHECK(ctx, cudaMemcpyAsync(out_ce, nat.p, M * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// constraint commitment (hook 3)
// ------------------------------------------------------------------------------------------------
// ce_cm: coset-major CE evaluations (ce x n) on device; fills m.polys (num_cols x n) and commits.
// E = 2 (quadratic extension): every composition column is an E column (v, 0); it is committed, opened and evaluated as the two
// base columns v and 0, so m.width = 2 * num_cols with the odd columns identically zero.
__global__ void k_spread_columns(const fe *__restrict__ src, size_t n, size_t ncols, fe *__restrict__ dst) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * ncols * n) return;
  size_t c = t / n, i = t - c * n;
  st_fe(dst + t, (c & 1) ? fe_zero() : ld_fe(src + (c >> 1) * n + i));
}
static int dev_constraint_commit(zkl_ctx *ctx, fe *ce_cm, size_t n, size_t ce, size_t num_cols, const zkl_options *o, committed_matrix &m,
                                 fe *overflow_flag, size_t E = 1, bool polys_only = false) {
  size_t M = n * ce;
  domain_tables *tM; ZKL_TRY(get_tables(ctx, M, &tM));
  m.width = num_cols * E; m.n = n; m.blowup = o->blowup; m.N = n * o->blowup; m.ext = E;
  dev_buf<fe> narrow;
  if (E == 2) ZKL_CHECK(ctx, narrow.alloc(num_cols * n));
  ZKL_CHECK(ctx, m.polys.alloc(m.width * n));
  ZKL_TRY(dev_interpolate(ctx, ce_cm, ce_cm, ce, n));  // per-coset inverse NTT in place
  auto it = ctx->opow.find(n | ((size_t)1 << 62));
  if (it == ctx->opow.end()) {
    fe *p; ZKL_CHECK(ctx, cudaMalloc((void **)&p, n * sizeof(fe)));
    LAUNCH(ctx, k_fill_powers, ceil_div(n, 256), 256, 0, fe_inv(fe_from_u64(3)), p, n);
    it = ctx->opow.emplace(n | ((size_t)1 << 62), p).first;
  }
  fe ginv_n = fe_pow(fe_inv(fe_from_u64(3)), n);
  LAUNCH(ctx, k_composition_finish, ceil_div(n, 128), 128, 0, ce_cm, (int)n, (int)ce, (int)num_cols, tM->wpow, M, it->second, ginv_n,
         fe_inv(fe_from_u64(ce)), E == 2 ? narrow.p : m.polys.p, overflow_flag);
  if (E == 2) LAUNCH(ctx, k_spread_columns, ceil_div(2 * num_cols * n, 256), 256, 0, narrow.p, n, num_cols, m.polys.p);
  if (polys_only) return 0;
  return commit_matrix(ctx, m, o);
}

extern "C" int zkl_constraint_commit(zkl_ctx *ctx, const zkl_fe *ce_evals, size_t n, size_t ce_blowup, size_t num_cols, const zkl_options *o,
                                     uint8_t out_root[32], zkl_fe *out_poly, zkl_constraint_commitment **out_handle) {
  ZKL_TRY(need_suite(ctx));
  ZKL_TRY(check_opts(ctx, o, n));
  if (num_cols > ce_blowup || ce_blowup > 16) { ctx->set_error("num_cols must be <= ce_blowup <= 16"); return 1; }
  size_t M = n * ce_blowup;
  zkl_constraint_commitment *h = new zkl_constraint_commitment();
  dev_buf<fe> nat, cm, flag;
  int rc = 1;
  do {
    if (nat.alloc(M) != cudaSuccess || cm.alloc(M) != cudaSuccess || flag.alloc(1) != cudaSuccess) { ctx->set_error("cudaMalloc failed"); break; }
    if (cudaMemcpyAsync(nat.p, ce_evals, M * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ctx->set_error("H2D failed"); break; }
    cudaMemsetAsync(flag.p, 0, sizeof(fe), ctx->stream);
    k_ce_to_natural<<<ceil_div(M, 256), 256, 0, ctx->stream>>>(nat.p, cm.p, (int)n, (int)ce_blowup, 0); ctx->launches++;
    if ((rc = dev_constraint_commit(ctx, cm.p, n, ce_blowup, num_cols, o, h->m, flag.p))) break;
    rc = 1;
    fe root;
    if (cudaMemcpyAsync(&root, h->m.nodes.p + 1, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ctx->set_error("D2H failed"); break; }
    if (out_poly && cudaMemcpyAsync(out_poly, h->m.polys.p, num_cols * n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ctx->set_error("D2H failed"); break; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->set_error(std::string("constraint commit: ") + cudaGetErrorString(cudaGetLastError())); break; }
    digest_from_fe_host(root, out_root);
    rc = 0;
  } while (0);
  if (rc) { delete h; return rc; }
  *out_handle = h;
  return 0;
}
extern "C" void zkl_constraint_commitment_free(zkl_ctx *ctx, zkl_constraint_commitment *h) {
  if (!h) return;
  enter(ctx); cudaStreamSynchronize(ctx->stream);
  delete h;
}

// ------------------------------------------------------------------------------------------------
// winterfell::Proof serialization (winter-air / winter-utils byte formats [WF-mem]) — byte shuffling only
// ------------------------------------------------------------------------------------------------
struct byte_writer {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { u8((uint8_t)v); u8((uint8_t)(v >> 8)); }
  void u32(uint32_t v) { for (int i = 0; i < 4; i++) u8((uint8_t)(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; i++) u8((uint8_t)(v >> (8 * i))); }
  void bytes(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
  void usize(uint64_t value) {  // winter-utils write_usize (vint64)
    int zeros = value == 0 ? 64 : __builtin_clzll(value);
    int len = (zeros > 0 ? zeros - 1 : 0) / 7; if (len > 8) len = 8;
    int length = 9 - len;
    if (length == 9) { u8(0); u64(value); }
    else { uint64_t enc = ((value << 1) | 1) << (length - 1); for (int i = 0; i < length; i++) u8((uint8_t)(enc >> (8 * i))); }
  }
  void digest(const fe &x) { bytes(x.l, 16); for (int i = 0; i < 16; i++) u8(0); }
  void felt(const fe &x) { bytes(x.l, 16); }
};
static void write_batch_proof_bytes(byte_writer &w, int depth, const std::vector<std::vector<fe>> &lists) {
  w.u8((uint8_t)depth); w.usize(lists.size());
  for (auto &l : lists) { w.usize(l.size()); for (auto &d : l) w.digest(d); }
}
static void write_context_bytes(byte_writer &w, size_t W, size_t n, const zkl_options &o, size_t num_constraints) {
  static const uint8_t modulus_le[16] = {0x01, 0x00, 0x00, 0x00, 0x00, 0xd3, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff};
  w.u8((uint8_t)W); w.u8(0); w.u8((uint8_t)ilog2_sz(n)); w.u16(0);
  w.u8(16); w.bytes(modulus_le, 16);
  w.usize(num_constraints);
  w.u8((uint8_t)o.num_queries); w.u8((uint8_t)o.blowup); w.u8((uint8_t)o.grinding); w.u8(o.field_extension == 2 ? 2 : 1);
  w.u8((uint8_t)o.fri_folding); w.u8((uint8_t)o.fri_remainder_max_degree); w.u8(0); w.u8(0);
  w.u8((uint8_t)o.num_partitions); w.u8((uint8_t)o.hash_rate);
}
static size_t fri_num_layers(size_t domain, const zkl_options &o) {  // FriOptions::num_fri_layers [WF-mem]
  size_t max_rem = (size_t)(o.fri_remainder_max_degree + 1) * o.blowup, r = 0;
  while (domain > max_rem) { domain /= o.fri_folding; r++; }
  return r;
}
static size_t rem_cap_of(const zkl_options &o) { return (size_t)o.fri_remainder_max_degree + 1; }
static std::vector<uint32_t> fold_positions_host(const std::vector<uint32_t> &pos, size_t src_domain) {  // agg/child.rs:1072-1099
  size_t tgt = src_domain / 2;
  std::vector<uint32_t> r;
  for (uint32_t p : pos) { uint32_t i = (uint32_t)(p % tgt); if (std::find(r.begin(), r.end(), i) == r.end()) r.push_back(i); }
  return r;
}

// internal OOD layout [T(z): W | H(z): C | T(zg): W | H(zg): C] (C already counts the zero halves) -> the wire layout of the
// quadratic extension: trace values widened to (a, 0)
__global__ void k_ood_wire(const fe *__restrict__ ood, int W, int C, fe *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = 2 * W + C;
  if (t >= 2 * half) return;
  int h = t / half, r = t - h * half;
  fe v;
  if (r < 2 * W) v = (r & 1) ? fe_zero() : ld_fe(ood + h * (W + C) + (r >> 1));
  else v = ld_fe(ood + h * (W + C) + W + (r - 2 * W));
  st_fe(out + t, v);
}
// FRI leaves of an extension-field layer: leaf i = k + B i' hashes (ev[k][i'], 0, ev[k][i' + nd/2], 0); out is column-major [4][half]
__global__ void k_fri_pairs_ext(const fe *__restrict__ ev, int nd, int B, fe *__restrict__ out) {
  size_t half_rows = (size_t)nd / 2, total = half_rows * B;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  size_t k = t / half_rows, ip = t - k * half_rows, leaf = k + (size_t)B * ip;
  st_fe(out + leaf, ld_fe(ev + k * nd + ip)); st_fe(out + total + leaf, fe_zero());
  st_fe(out + 2 * total + leaf, ld_fe(ev + k * nd + ip + half_rows)); st_fe(out + 3 * total + leaf, fe_zero());
}
// ---- column-sharded proving (include/zkl_b200.h: zkl_comm) ----
// Ownership. Trace COLUMNS: rank s owns the whole partitions [s P / R, (s + 1) P / R) (balanced: X's 14 partitions over 8 ranks are
// 6 x 2 + 2 x 1). LDE COSETS of everything that is not column-sharded (constraint evaluations, composition columns, DEEP values, FRI
// layers): rank s owns cosets [s B / R, (s + 1) B / R) — the coset-major layout makes a coset range one contiguous slab per column and
// FRI folding never leaves a coset. Merkle LEAVES: rank s builds the sub-tree over the natural rows [s N / R, (s + 1) N / R).
struct shard_view { const zkl_comm *comm; size_t first_col, local_width; };
struct shard_geo {
  const zkl_comm *comm = nullptr;
  int R = 1, rk = 0;
  size_t ps = 0, P = 0, ppr = 0;   // partition size, partitions, most partitions any rank owns
  size_t min_leaves = 0;           // trees and FRI layers below this many leaves are built on every rank instead of exchanged
  void parts_of(int s, size_t *first, size_t *count) const { *first = (size_t)s * P / R; *count = (size_t)(s + 1) * P / R - *first; }
  void cols_of(int s, size_t W, size_t *first, size_t *count) const {
    size_t f, c; parts_of(s, &f, &c);
    size_t lo = std::min(W, f * ps), hi = std::min(W, (f + c) * ps);
    *first = lo; *count = hi - lo;
  }
};
// send[d][c][j][i] = lde[c][(d * ncl + j) * step][i]: the CE cosets every destination rank evaluates, for this rank's columns (rows of
// columns beyond `wl` stay zero: every rank sends blocks of the same size)
__global__ void k_pack_cosets(const fe *__restrict__ lde, size_t N, int n, int step, int wl, int wmax, int ncl, int world, fe *__restrict__ send) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per_dest = (size_t)wmax * ncl * n;
  if (t >= per_dest * world) return;
  int d = (int)(t / per_dest); size_t u = t - (size_t)d * per_dest;
  int c = (int)(u / ((size_t)ncl * n)); size_t v = u - (size_t)c * ncl * n;
  int j = (int)(v / n), i = (int)(v - (size_t)j * n);
  st_fe(send + t, c < wl ? ld_fe(lde + (size_t)c * N + (size_t)((d * ncl + j) * step) * n + i) : fe_zero());
}
// out[t] = sum over s of in[s * count + t] (field addition): the DEEP partial sums of the ranks for the cosets this rank owns
__global__ void k_sum_slices(const fe *__restrict__ in, int slices, size_t count, fe *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  fe a = ld_fe(in + t);
  for (int s = 1; s < slices; s++) a = fe_add(a, ld_fe(in + (size_t)s * count + t));
  st_fe(out + t, a);
}
// gathered[s][k' + Bp i] (every rank's leaves of ITS Bp cosets, natural order inside the rank) -> natural order over all B cosets:
// out[k + B i] with k = s Bp + k'
__global__ void k_interleave_leaves(const fe *__restrict__ g, fe *__restrict__ out, int B, int Bp, size_t m) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)B * m) return;
  size_t k = p % B, i = p / B;
  st_fe(out + p, ld_fe(g + (k / Bp) * ((size_t)Bp * m) + (k % Bp) + (size_t)Bp * i));
}
// heap node g >= R of the whole tree out of the gathered sub-trees: level lvl = 2^floor(log2 g) holds lvl / R nodes of each rank
__global__ void k_scatter_subtrees(const fe *__restrict__ gathered, fe *__restrict__ nodes, size_t count, int R) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x + R;
  if (g >= count) return;
  size_t lvl = (size_t)1 << (63 - __clzll((unsigned long long)g)), j = g - lvl, per = lvl / R, s = j / per;
  st_fe(nodes + g, ld_fe(gathered + s * (count / R) + per + (j - s * per)));
}
static int comm_call(zkl_ctx *ctx, int rc, const char *what) {
  if (rc) { ctx->set_error(std::string("collective callback failed: ") + what); return 1; }
  return 0;
}
static int shard_geometry(zkl_ctx *ctx, const zkl_options *o, size_t W, const shard_view *sh, size_t ce, shard_geo &g) {
  g.comm = sh->comm; g.R = sh->comm->world; g.rk = sh->comm->rank;
  if (g.R < 1 || g.rk < 0 || g.rk >= g.R) { ctx->set_error("zkl_comm: bad rank / world"); return 1; }
  if (!sh->comm->all_gather || !sh->comm->all_to_all || !sh->comm->all_reduce_sum_u64) { ctx->set_error("zkl_comm: missing collective callback"); return 1; }
  if (ce % g.R || o->blowup % g.R) { ctx->set_error("the world size must divide the constraint-evaluation blowup (one or more CE cosets per rank)"); return 1; }
  size_t f, wl;
  if (zkl_shard_columns(o, W, g.rk, g.R, &f, &wl)) { ctx->set_error("column sharding needs at least one whole partition per rank (PartitionOptions with >= world partitions)"); return 1; }
  if (f != sh->first_col || wl != sh->local_width) { ctx->set_error("first_col / local_width do not match zkl_shard_columns for this rank"); return 1; }
  g.ps = partition_size_of(o, W); g.P = (W + g.ps - 1) / g.ps; g.ppr = (g.P + g.R - 1) / g.R;
  g.min_leaves = ctx->shard_min_leaves;
  return 0;
}
// MerkleTree::new over `count` leaves that every rank holds: each rank builds the sub-tree over its contiguous share, the sub-trees
// are all-gathered (count / R nodes per rank), scattered into heap order and topped with the log2(R) levels above them
static int sharded_merkle(zkl_ctx *ctx, const shard_geo &g, const fe *leaves, fe *nodes, size_t count) {
  const size_t per = count / g.R;
  if (g.R == 1 || count < g.min_leaves || per < 2) return dev_merkle(ctx, leaves, nodes, count);
  dev_buf<fe> sub, all;
  ZKL_CHECK(ctx, sub.alloc(per)); ZKL_CHECK(ctx, all.alloc(count));
  ZKL_CHECK(ctx, cudaMemsetAsync(sub.p, 0, sizeof(fe), ctx->stream));   // heap slot 0 is unused; it travels with the rest
  ZKL_TRY(dev_merkle(ctx, leaves + (size_t)g.rk * per, sub.p, per));
  ZKL_TRY(ctx_wait(ctx));
  ZKL_TRY(comm_call(ctx, g.comm->all_gather(g.comm->user, sub.p, all.p, per * sizeof(fe)), "all_gather(sub-trees)"));
  LAUNCH(ctx, k_scatter_subtrees, ceil_div(count - g.R, 256), 256, 0, all.p, nodes, count, g.R);
  ZKL_TRY(dev_merkle(ctx, nodes + g.R, nodes, (size_t)g.R));
  return ctx_wait(ctx);   // `all` is released when this returns: its last reader has run
}
// local: the leaves of the Bp = B / R cosets this rank owns, Bp * m of them in the rank's own natural order (k' + Bp i); result: all
// B * m leaves in natural order and the whole tree on every rank
static int shard_leaves_to_tree(zkl_ctx *ctx, const shard_geo &g, const fe *local, size_t B, size_t m, fe *leaves, fe *nodes) {
  const size_t Bp = B / g.R;
  dev_buf<fe> all; ZKL_CHECK(ctx, all.alloc(B * m));
  ZKL_TRY(ctx_wait(ctx));
  ZKL_TRY(comm_call(ctx, g.comm->all_gather(g.comm->user, local, all.p, Bp * m * sizeof(fe)), "all_gather(leaves)"));
  LAUNCH(ctx, k_interleave_leaves, ceil_div(B * m, 256), 256, 0, all.p, leaves, (int)B, (int)Bp, m);
  ZKL_TRY(sharded_merkle(ctx, g, leaves, nodes, B * m));
  return ctx_wait(ctx);
}
static const char *k_stage_names[12] = {"coin_seed", "trace_lde_commit", "constraint_eval", "constraint_commit", "ood", "deep",
                                        "fri", "grind", "queries_serialize", "kernel:k_lde(trace)", "kernel:k_hash_rows(trace)",
                                        "host:air_build"};
extern "C" const char *zkl_stage_name(int s) { return (s >= 0 && s < 12) ? k_stage_names[s] : ""; }

// ------------------------------------------------------------------------------------------------
// whole-segment prover: winterfell::Prover::prove, stages 1-11, everything on the device
// ------------------------------------------------------------------------------------------------
static int prove_segment_impl(zkl_ctx *ctx, const fe *trace_dev, size_t W, size_t n, const zkl_air_params *air, const zkl_fe *seed_elems,
                              size_t num_seed, const zkl_options *o, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary,
                              const shard_view *sh = nullptr, const sched_assertions *sched = nullptr) {
  ZKL_TRY(need_suite(ctx));
  ZKL_TRY(check_opts(ctx, o, n));
  // column-sharded mode: trace_dev holds this rank's Wl columns [first_col, first_col + Wl) of the W-column trace
  const size_t Wl = sh ? sh->local_width : W, col0 = sh ? sh->first_col : 0;
  const int R = sh ? sh->comm->world : 1, rk = sh ? sh->comm->rank : 0;
  shard_geo sg;
  if (sh) {
    ZKL_TRY(shard_geometry(ctx, o, W, sh, air->ce_blowup, sg));
    if (o->field_extension == 2) { ctx->set_error("column sharding is for step proofs (FieldExtension::None)"); return 1; }
  }
  const size_t Bp = o->blowup / (size_t)R, k0 = (size_t)rk * Bp;   // the LDE cosets this rank owns: [k0, k0 + Bp)
  auto wait_stream = [&]() -> int { return ctx_wait(ctx); };
  // ZKL_SHARD_TRACE=1: rank 0 prints the host wall time between the marked points of the column-sharded flow (each mark drains the stream)
  const bool shard_trace = sh && rk == 0 && getenv("ZKL_SHARD_TRACE");
  auto trace_t0 = std::chrono::steady_clock::now();
  auto mark = [&](const char *what) {
    if (!shard_trace) return;
    ctx_wait(ctx);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[shard] %-44s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - trace_t0).count());
    trace_t0 = t1;
  };
  if (air->cols.width != W) { ctx->set_error("air width does not match the trace"); return 1; }
  if (o->num_queries == 0 || o->num_queries > 255) { ctx->set_error("num_queries must be in [1,255]"); return 1; }
  // E = 2: the quadratic extension. Every drawn E element is (a, 0) with this hasher (zkl_options.field_extension), so the arithmetic
  // below is the base-field arithmetic of a step proof; C E-columns of the composition polynomial travel as Cw = 2 C base columns
  // (odd ones zero), and wherever Winterfell hashes or serializes E elements the pairs (a, 0) are hashed / written.
  const size_t E = o->field_extension == 2 ? 2 : 1;
  const size_t B = o->blowup, N = n * B, ce = air->ce_blowup, M = n * ce, C = (size_t)air->num_composition_cols * E, Wt = W * E;
  const size_t ncc = (size_t)air->num_transition + air->num_assertions;
  const size_t nl = fri_num_layers(N, *o);
  if (nl + 1 > 40) { ctx->set_error("too many FRI layers"); return 1; }
  while (ctx->ev.size() < 16) { cudaEvent_t e; ZKL_CHECK(ctx, cudaEventCreate(&e)); ctx->ev.push_back(e); }
  // pinned staging for everything this proof moves between host and device besides the trace: AIR parameters, seed,
  // the opening plan (one pointer + one index per opened element) and the opened elements themselves
  const size_t logN = (size_t)ilog2_sz(N);
  const size_t max_open = (size_t)o->num_queries * (W + C + 2 * logN + nl * (2 + logN)) + 64;
  const size_t stage_as = sched ? 4096 + (sched->tmpl.size() + 2 * sched->extras.size() + 8) * 64 : ((size_t)air->num_assertions + 2) * (sizeof(assertion_dev) + 4);
  const size_t stage_need = sizeof(air_dev) + stage_as + 256 * 4 + num_seed * sizeof(fe) +
                            max_open * (sizeof(fe *) + 4 + sizeof(fe)) + (4 * (W + C) + nl + rem_cap_of(*o) + 64) * sizeof(fe) + 8192;
  ZKL_TRY(stage_reserve(ctx, stage_need));
  domain_tables *tN; ZKL_TRY(get_tables(ctx, N, &tN));
  fe *opow; ZKL_TRY(get_offset_powers(ctx, n, &opow)); (void)opow;
  cudaStream_t st = ctx->stream;
  dev_buf<fe> coin;
  const bool quad = use_quad(ctx);   // low-latency sponges on 64-thread groups (kernels_quad.cuh), else 16-thread groups
  // coin.reseed(digest) followed by `draws` draws, counter restarting at 1 (DefaultRandomCoin [WF-mem]); the digest is
  // also copied to *copy_to when given
  auto reseed_draw = [&](const fe *digest, fe *copy_to, size_t draws, fe *out) -> int {
    const size_t fused = draws <= ZKL_GRP_MAX_GROUPS ? draws : 0;   // few draws ride in the reseed launch
    const unsigned groups = (unsigned)std::max<size_t>(fused, 1);
    if (quad) LAUNCH(ctx, k_reseed_draw<ZKL_QUAD_GROUP>, 1, groups * ZKL_QUAD_GROUP, 0, coin.p, digest, copy_to, (int)fused, out);
    else LAUNCH(ctx, k_reseed_draw<ZKL_COOP_GROUP>, 1, groups * ZKL_COOP_GROUP, 0, coin.p, digest, copy_to, (int)fused, out);
    if (draws > fused) {
      if (draws >= bulk_min(ctx) && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_coin_draw, draws, coin.p, (uint64_t)0, draws, out);
      else if (quad) LAUNCH_QUAD(ctx, k_quad_coin_draw, draws, coin.p, (uint64_t)0, (int)draws, out);
      else LAUNCH_COOP(ctx, k_coin_draw, draws, coin.p, (uint64_t)0, (int)draws, out);
    }
    return 0;
  };
  int evi = 0;
#define STAGE_MARK() ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[evi++], st))
  STAGE_MARK();

  // ---- 1. public coin seed (agg/fs.rs:69-73)
  dev_buf<fe> seed_in, tmp;
  ZKL_CHECK(ctx, seed_in.alloc(num_seed)); ZKL_CHECK(ctx, coin.alloc(1)); ZKL_CHECK(ctx, tmp.alloc(8));
  {
    void *sh = stage_alloc(ctx, num_seed * sizeof(fe));
    memcpy(sh, seed_elems, num_seed * sizeof(fe));
    ZKL_CHECK(ctx, cudaMemcpyAsync(seed_in.p, sh, num_seed * sizeof(fe), cudaMemcpyHostToDevice, st));
  }
  if (quad) LAUNCH(ctx, k_quad_coin_init, 1, ZKL_QUAD_GROUP, 0, seed_in.p, (int)num_seed, coin.p);
  else LAUNCH(ctx, k_coin_init, 1, ZKL_COOP_GROUP, 0, seed_in.p, (int)num_seed, coin.p);
  STAGE_MARK();

  // ---- 2. trace LDE + commitment (hook new_trace_lde), reseed with the root
  zkl_trace_lde tl;
  air_device_state ast;
  if (!sh) ZKL_TRY(trace_commit_dev(ctx, trace_dev, W, n, o, &tl, &ctx->ev[12]));
  else {
    // local: iNTT, coset LDE and partition digests of the owned columns. Exchange (1): 16 bytes per (row, partition) instead of the
    // rows — an all-to-all that hands every rank all P digests of ITS N / R rows; it merges them into leaves, the leaves are
    // all-gathered and the tree is built from per-rank sub-trees (sharded_merkle)
    committed_matrix &m = tl.m;
    m.width = Wl; m.n = n; m.blowup = B; m.N = N;
    ZKL_CHECK(ctx, m.polys.alloc(Wl * n)); ZKL_CHECK(ctx, m.lde.alloc(Wl * N)); ZKL_CHECK(ctx, m.leaves.alloc(N)); ZKL_CHECK(ctx, m.nodes.alloc(N));
    mark("start, allocations");
    ZKL_TRY(dev_interpolate(ctx, trace_dev, m.polys.p, Wl, n));
    mark("trace: interpolate");
    ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[12], st));
    ZKL_TRY(dev_lde(ctx, m.polys.p, m.lde.p, Wl, n, B));
    ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[13], st));
    mark("trace: coset LDE");
    size_t pf, pc; sg.parts_of(rk, &pf, &pc);
    const size_t per = N / R;
    dev_buf<fe> mine, send, recv, parts, leaves_loc;
    ZKL_CHECK(ctx, mine.alloc(pc * N)); ZKL_CHECK(ctx, send.alloc((size_t)R * sg.ppr * per)); ZKL_CHECK(ctx, recv.alloc((size_t)R * sg.ppr * per));
    ZKL_CHECK(ctx, parts.alloc(sg.P * per)); ZKL_CHECK(ctx, leaves_loc.alloc(per));
    ZKL_TRY(dev_hash_partitions(ctx, m.lde.p, Wl, n, B, sg.ps, mine.p));
    ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[14], st));
    // the host orders and stages the assertion list (millions of entries for a large segment) while the device extends and hashes
    if (!shard_trace) ZKL_TRY(upload_air(ctx, air, n, ast, sched));
    mark("trace: partition digests");
    if (pc < sg.ppr) ZKL_CHECK(ctx, cudaMemsetAsync(send.p, 0, (size_t)R * sg.ppr * per * sizeof(fe), st));   // padded slots travel too
    for (int d = 0; d < R; d++) {   // send[d][q][j] = mine[q][d per + j]
      if (N * sizeof(fe) <= 0x7FFFFFFFull)   // a strided copy while the source pitch fits cudaMemcpy2D's limit (2^22 rows at blowup 16 do)
        ZKL_CHECK(ctx, cudaMemcpy2DAsync(send.p + (size_t)d * sg.ppr * per, per * sizeof(fe), mine.p + (size_t)d * per, N * sizeof(fe), per * sizeof(fe), pc,
                                         cudaMemcpyDeviceToDevice, st));
      else
        for (size_t q = 0; q < pc; q++)
          ZKL_CHECK(ctx, cudaMemcpyAsync(send.p + ((size_t)d * sg.ppr + q) * per, mine.p + q * N + (size_t)d * per, per * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    }
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_to_all(sg.comm->user, send.p, recv.p, sg.ppr * per * sizeof(fe)), "all_to_all(partition digests)"));
    mark("trace: pack + all_to_all(digests)");
    for (int s2 = 0; s2 < R; s2++) {   // parts[q][j], q in global partition order
      size_t f2, c2; sg.parts_of(s2, &f2, &c2);
      ZKL_CHECK(ctx, cudaMemcpyAsync(parts.p + f2 * per, recv.p + (size_t)s2 * sg.ppr * per, c2 * per * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    }
    ZKL_TRY(dev_merge_partitions(ctx, parts.p, per, per, sg.P, leaves_loc.p));
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_gather(sg.comm->user, leaves_loc.p, m.leaves.p, per * sizeof(fe)), "all_gather(trace leaves)"));
    mark("trace: merge_many + all_gather(leaves)");
    ZKL_TRY(sharded_merkle(ctx, sg, m.leaves.p, m.nodes.p, N));
    ZKL_TRY(wait_stream());
    mark("trace: tree (sub-trees + all_gather)");
  }
  STAGE_MARK();

  // ---- 4/5. composition coefficients (transition then boundary) and constraint evaluation
  if (!sh || shard_trace) ZKL_TRY(upload_air(ctx, air, n, ast, sched));
  dev_buf<fe> cc, ce_ev, flag;
  // column-sharded prover with millions of coefficients (one per assertion): every rank draws 1 / R of them, all-gathered
  const size_t cc_per = (ncc + (size_t)R - 1) / (size_t)R;
  const bool cc_sharded = sh && R > 1 && ncc >= sg.min_leaves && ctx->hash_path == 0 && cc_per >= bulk_min(ctx);
  ZKL_CHECK(ctx, cc.alloc(cc_sharded ? cc_per * (size_t)R : ncc)); ZKL_CHECK(ctx, ce_ev.alloc(M)); ZKL_CHECK(ctx, flag.alloc(1));
  ZKL_CHECK(ctx, cudaMemsetAsync(flag.p, 0, sizeof(fe), st));
  mark("air upload (host: assertions sorted, staged)");
  if (!cc_sharded) ZKL_TRY(reseed_draw(tl.m.nodes.p + 1, nullptr, ncc, cc.p));   // reseed with the trace root, draw the composition coefficients
  else {
    dev_buf<fe> mine; ZKL_CHECK(ctx, mine.alloc(cc_per));
    ZKL_TRY(reseed_draw(tl.m.nodes.p + 1, nullptr, 0, nullptr));
    LAUNCH_TC(ctx, k_tc_coin_draw, cc_per, coin.p, (uint64_t)((size_t)rk * cc_per), cc_per, mine.p);   // draws rk cc_per + 1 ... (past ncc: unused)
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_gather(sg.comm->user, mine.p, cc.p, cc_per * sizeof(fe)), "all_gather(composition coefficients)"));
  }
  mark("draw composition coefficients");
  if (!sh) ZKL_TRY(dev_eval_constraints(ctx, tl.m, air, ast, cc.p, ce_ev.p));
  else {
    // exchange (2): every rank evaluates ce / R constraint-evaluation cosets on FULL-width rows — the coset-major layout makes a coset
    // the natural unit: all-to-all of [columns of the sender][cosets of the receiver][n], then all-gather of the evaluations
    const size_t ncl = ce / R, wmax = sg.ppr * sg.ps, per_pair = wmax * ncl * n;
    dev_buf<fe> send, recv, rows, ce_all;
    ZKL_CHECK(ctx, send.alloc(per_pair * R)); ZKL_CHECK(ctx, recv.alloc(per_pair * R)); ZKL_CHECK(ctx, rows.alloc(W * ncl * n)); ZKL_CHECK(ctx, ce_all.alloc(M));
    LAUNCH(ctx, k_pack_cosets, ceil_div(per_pair * R, 256), 256, 0, tl.m.lde.p, N, (int)n, (int)(B / ce), (int)Wl, (int)wmax, (int)ncl, R, send.p);
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sh->comm->all_to_all(sh->comm->user, send.p, recv.p, per_pair * sizeof(fe)), "all_to_all(CE cosets)"));
    mark("ce: pack + all_to_all(CE cosets)");
    for (int s2 = 0; s2 < R; s2++) {
      size_t f, cnt; sg.cols_of(s2, W, &f, &cnt);
      if (cnt) ZKL_CHECK(ctx, cudaMemcpyAsync(rows.p + f * ncl * n, recv.p + (size_t)s2 * per_pair, cnt * ncl * n * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    }
    ce_rows cr{rows.p, (int)(rk * ncl), (int)ncl};
    ZKL_TRY(dev_eval_constraints(ctx, tl.m, air, ast, cc.p, ce_ev.p, &cr));
    ZKL_TRY(wait_stream());
    mark("ce: boundary polys + constraint kernel");
    ZKL_TRY(comm_call(ctx, sh->comm->all_gather(sh->comm->user, ce_ev.p + (size_t)rk * ncl * n, ce_all.p, ncl * n * sizeof(fe)), "all_gather(CE evaluations)"));
    ZKL_CHECK(ctx, cudaMemcpyAsync(ce_ev.p, ce_all.p, M * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    ZKL_TRY(wait_stream());   // ce_all, send, recv die at the end of this block: their frees are stream-ordered, the copies above are done
    mark("ce: all_gather(evaluations)");
  }
  STAGE_MARK();

  // ---- 6. constraint commitment (hook build_constraint_commitment), reseed
  zkl_constraint_commitment cmt;
  if (!sh) ZKL_TRY(dev_constraint_commit(ctx, ce_ev.p, n, ce, air->num_composition_cols, o, cmt.m, flag.p, E));
  else {
    // the composition columns are few: coefficients and LDE on every rank; the row hashes of the cosets this rank owns, leaves
    // all-gathered, tree from sub-trees
    committed_matrix &m = cmt.m;
    ZKL_TRY(dev_constraint_commit(ctx, ce_ev.p, n, ce, air->num_composition_cols, o, m, flag.p, E, true));
    // m.lde holds THIS RANK'S cosets only, compact: [column][Bp][n]
    ZKL_CHECK(ctx, m.lde.alloc(m.width * Bp * n)); ZKL_CHECK(ctx, m.leaves.alloc(N)); ZKL_CHECK(ctx, m.nodes.alloc(N));
    ZKL_TRY(dev_lde_cosets(ctx, m.polys.p, m.lde.p, m.width, n, B, k0, Bp));
    dev_buf<fe> loc; ZKL_CHECK(ctx, loc.alloc(Bp * n));
    ZKL_TRY(dev_hash_rows(ctx, m.lde.p, m.width, n, Bp, o, loc.p, m.scratch, m.ext));
    ZKL_TRY(shard_leaves_to_tree(ctx, sg, loc.p, B, n, m.leaves.p, m.nodes.p));
    mark("constraint commitment");
  }
  STAGE_MARK();

  // ---- 7. OOD point z, evaluations at z and z*g, reseed with hash_elements(cur(z)||quot(z)||next(zg)||quot(zg))
  dev_buf<fe> z, zpow, ood, ood_digest;
  ZKL_CHECK(ctx, z.alloc(1)); ZKL_CHECK(ctx, zpow.alloc(2 * n)); ZKL_CHECK(ctx, ood.alloc(2 * (W + C))); ZKL_CHECK(ctx, ood_digest.alloc(1));
  ZKL_TRY(reseed_draw(cmt.m.nodes.p + 1, nullptr, 1, z.p));   // reseed with the constraint root, draw z
  fe g_n = host_root_of_unity(ilog2_sz(n));
  LAUNCH(ctx, k_powers_dev, ceil_div(n, 128), 128, 0, z.p, fe_one(), zpow.p, (int)n);
  LAUNCH(ctx, k_powers_dev, ceil_div(n, 128), 128, 0, z.p, g_n, zpow.p + n, (int)n);
  if (sh) ZKL_CHECK(ctx, cudaMemsetAsync(ood.p, 0, 2 * (W + C) * sizeof(fe), st));
  // T(z), T(z g) of the (owned) trace columns and H(z), H(z g): both points in one pass over the coefficients
  const size_t ev_chunks = (n + POLY_EVAL_CHUNK - 1) / POLY_EVAL_CHUNK;
  dev_buf<fe> ev_part;
  if (ev_chunks > 1) ZKL_CHECK(ctx, ev_part.alloc(2 * ev_chunks * std::max(Wl, C)));
  auto eval_two_points = [&](const fe *polys, size_t count, fe *at_z, fe *at_zg) -> int {
    LAUNCH(ctx, k_poly_eval2, dim3((unsigned)ev_chunks, (unsigned)count), 256, 0, polys, (int)n, zpow.p, ev_part.p, at_z, at_zg);
    if (ev_chunks > 1) LAUNCH(ctx, k_poly_eval2_finish, ceil_div(2 * count, 128), 128, 0, ev_part.p, (int)ev_chunks, (int)count, at_z, at_zg);
    return 0;
  };
  ZKL_TRY(eval_two_points(tl.m.polys.p, Wl, ood.p + col0, ood.p + W + C + col0));
  if (sh) {   // exchange (3a): every rank evaluated its own columns; the supports are disjoint, so a u64 sum assembles the frame exactly
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sh->comm->all_reduce_sum_u64(sh->comm->user, ood.p, 2 * W), "all_reduce(OOD trace, z)"));
    ZKL_TRY(comm_call(ctx, sh->comm->all_reduce_sum_u64(sh->comm->user, ood.p + W + C, 2 * W), "all_reduce(OOD trace, zg)"));
  }
  ZKL_TRY(eval_two_points(cmt.m.polys.p, C, ood.p + W, ood.p + 2 * W + C));
  // what is hashed and sent: the frame as E elements, [T(z) | H(z) | T(zg) | H(zg)] with every trace value widened to (a, 0)
  dev_buf<fe> ood_wide;
  const fe *ood_w = ood.p;
  if (E == 2) {
    ZKL_CHECK(ctx, ood_wide.alloc(2 * (Wt + C)));
    LAUNCH(ctx, k_ood_wire, ceil_div(2 * (Wt + C), 128), 128, 0, ood.p, (int)W, (int)C, ood_wide.p);
    ood_w = ood_wide.p;
  }
  if (quad) LAUNCH(ctx, k_quad_hash_elements_single, 1, ZKL_QUAD_GROUP, 0, ood_w, (int)(2 * (Wt + C)), ood_digest.p);
  else LAUNCH(ctx, k_hash_elements_single, 1, ZKL_COOP_GROUP, 0, ood_w, (int)(2 * (Wt + C)), ood_digest.p);
  STAGE_MARK();

  // ---- 8. DEEP coefficients + composition over the LDE domain
  dev_buf<fe> dc, az, deep;
  ZKL_CHECK(ctx, dc.alloc(W + C)); ZKL_CHECK(ctx, az.alloc(2)); ZKL_CHECK(ctx, deep.alloc(N));
  if (E == 1) ZKL_TRY(reseed_draw(ood_digest.p, nullptr, W + C, dc.p));   // reseed with the OOD digest, draw the DEEP coefficients
  else {   // one coefficient per E column: the zero halves of the composition columns get none
    dev_buf<fe> drawn; ZKL_CHECK(ctx, drawn.alloc(W + C / 2));
    ZKL_TRY(reseed_draw(ood_digest.p, nullptr, W + C / 2, drawn.p));
    ZKL_CHECK(ctx, cudaMemcpyAsync(dc.p, drawn.p, W * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    LAUNCH(ctx, k_spread_columns, ceil_div(C, 256), 256, 0, drawn.p + W, (size_t)1, C / 2, dc.p + W);
  }
  mark("ood frame");
  dev_buf<fe> deep_inv; ZKL_CHECK(ctx, deep_inv.alloc(N));
  LAUNCH(ctx, k_deep_inverses, ceil_div(N, 256 * DEEP_INV_PER_THREAD), 256, 0, N, (int)n, (int)B, z.p, g_n, tN->wpow, deep_inv.p);
  if (!sh) {
    LAUNCH(ctx, k_deep_consts, 1, 32, 0, ood.p, dc.p, (int)W, (int)C, az.p);
    LAUNCH(ctx, k_deep, ceil_div(N, 256), 256, 0, tl.m.lde.p, cmt.m.lde.p, N, (int)n, (int)B, (int)W, (int)C, dc.p, az.p, z.p, g_n, tN->wpow, deep_inv.p, deep.p,
           (size_t)0, N, 0, N, (size_t)0);
  } else {
    // The DEEP composition is linear in the columns. Every rank sums over ITS trace columns on the whole domain; exchange (3b) is a
    // reduce-scatter by coset range (an all-to-all of the partial sums and a field addition of the R slices): the rank ends up with
    // the values on the cosets it owns — where FRI continues — and adds the composition columns' share there.
    dev_buf<fe> dcl, oodl, recv, oodc, azc;
    ZKL_CHECK(ctx, dcl.alloc(Wl)); ZKL_CHECK(ctx, oodl.alloc(2 * Wl)); ZKL_CHECK(ctx, recv.alloc(N)); ZKL_CHECK(ctx, oodc.alloc(2 * C)); ZKL_CHECK(ctx, azc.alloc(2));
    ZKL_CHECK(ctx, cudaMemcpyAsync(dcl.p, dc.p + col0, Wl * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(oodl.p, ood.p + col0, Wl * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(oodl.p + Wl, ood.p + W + C + col0, Wl * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(oodc.p, ood.p + W, C * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(oodc.p + C, ood.p + 2 * W + C, C * sizeof(fe), cudaMemcpyDeviceToDevice, st));
    LAUNCH(ctx, k_deep_consts, 1, 32, 0, oodl.p, dcl.p, (int)Wl, 0, az.p);
    LAUNCH(ctx, k_deep_consts, 1, 32, 0, oodc.p, dc.p + W, 0, (int)C, azc.p);
    LAUNCH(ctx, k_deep, ceil_div(N, 256), 256, 0, tl.m.lde.p, (const fe *)nullptr, N, (int)n, (int)B, (int)Wl, 0, dcl.p, az.p, z.p, g_n, tN->wpow, deep_inv.p, deep.p,
           (size_t)0, N, 0, N, (size_t)0);
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_to_all(sg.comm->user, deep.p, recv.p, Bp * n * sizeof(fe)), "all_to_all(DEEP partial sums)"));
    LAUNCH(ctx, k_sum_slices, ceil_div(Bp * n, 256), 256, 0, recv.p, R, Bp * n, deep.p);   // deep[0 .. Bp n): this rank's cosets from here on
    LAUNCH(ctx, k_deep, ceil_div(Bp * n, 256), 256, 0, (const fe *)nullptr, cmt.m.lde.p, N, (int)n, (int)B, 0, (int)C, dc.p + W, azc.p, z.p, g_n, tN->wpow, deep_inv.p,
           deep.p, k0 * n, Bp * n, 1, Bp * n, k0 * n);
    ZKL_TRY(wait_stream());
    mark("deep (partial sums, all_to_all, add)");
  }
  STAGE_MARK();

  // ---- 9. FRI layers
  std::vector<dev_buf<fe>> f_ev(nl + 1), f_leaves(nl), f_nodes(nl);
  std::vector<const fe *> layer_ev(nl + 1, nullptr);   // where layer d's evaluations are (for the openings)
  std::vector<char> f_sharded(nl + 1, 0);   // layer d's evaluations live as the Bp cosets this rank owns (column-sharded prover)
  dev_buf<fe> alphas, fri_roots, rem_rev;
  ZKL_CHECK(ctx, alphas.alloc(nl + 1)); ZKL_CHECK(ctx, fri_roots.alloc(nl + 1));
  fe g_inv_half = fe_inv(fe_from_u64(6));
  const fe *cur_ev = deep.p;
  size_t nd = n;
  bool ev_sharded = sh != nullptr;   // deep.p holds [Bp][n] then
  dev_buf<fe> ev_full;
  auto unshard_ev = [&]() -> int {   // all-gather the cosets: [R][Bp][nd] IS the coset-major layer
    ZKL_CHECK(ctx, ev_full.alloc(B * nd));
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_gather(sg.comm->user, cur_ev, ev_full.p, Bp * nd * sizeof(fe)), "all_gather(FRI layer)"));
    cur_ev = ev_full.p; ev_sharded = false;
    return 0;
  };
  for (size_t d = 0; d < nl; d++) {
    size_t half = nd * B / 2;
    if (ev_sharded && (half < sg.min_leaves || half / R < 2)) ZKL_TRY(unshard_ev());
    f_sharded[d] = ev_sharded; layer_ev[d] = cur_ev;
    const size_t Bd = ev_sharded ? Bp : B, half_d = nd * Bd / 2;   // cosets and leaves this rank works on
    ZKL_CHECK(ctx, f_leaves[d].alloc(half)); ZKL_CHECK(ctx, f_nodes[d].alloc(half)); ZKL_CHECK(ctx, f_ev[d + 1].alloc(half_d));
    dev_buf<fe> loc;
    fe *lv = f_leaves[d].p;
    if (ev_sharded) { ZKL_CHECK(ctx, loc.alloc(half_d)); lv = loc.p; }
    if (E == 2) {   // a leaf hashes two E elements: (v0, 0, v1, 0) — laid out as a 4-column matrix in natural leaf order
      dev_buf<fe> pairs; ZKL_CHECK(ctx, pairs.alloc(4 * half));
      LAUNCH(ctx, k_fri_pairs_ext, ceil_div(half, 256), 256, 0, cur_ev, (int)nd, (int)B, pairs.p);
      ZKL_TRY(dev_hash_partitions(ctx, pairs.p, 4, half, 1, 4, f_leaves[d].p));
    }
    else if (half_d >= bulk_min(ctx) && ctx->hash_path == 0) LAUNCH_TC(ctx, k_tc_fri_leaves, half_d, cur_ev, (int)nd, (int)Bd, lv);
    else if (quad) LAUNCH_QUAD(ctx, k_quad_fri_leaves, half_d, cur_ev, (int)nd, (int)Bd, lv);
    else LAUNCH_COOP(ctx, k_fri_leaves, half_d, cur_ev, (int)nd, (int)Bd, lv);
    if (ev_sharded) ZKL_TRY(shard_leaves_to_tree(ctx, sg, loc.p, B, nd / 2, f_leaves[d].p, f_nodes[d].p));
    else ZKL_TRY(dev_merkle(ctx, f_leaves[d].p, f_nodes[d].p, half));
    ZKL_TRY(reseed_draw(f_nodes[d].p + 1, fri_roots.p + d, 1, alphas.p + d));   // commit the layer root, draw alpha
    LAUNCH(ctx, k_fri_fold, ceil_div(half_d, 256), 256, 0, cur_ev, (int)nd, (int)Bd, (int)d, alphas.p + d, tN->wpow, N, g_inv_half, f_ev[d + 1].p,
           ev_sharded ? (int)k0 : 0, (int)B);
    cur_ev = f_ev[d + 1].p;
    nd /= 2;
  }
  dev_buf<fe> ev_last;
  if (ev_sharded) {   // the remainder is interpolated from the whole last layer
    ZKL_CHECK(ctx, ev_last.alloc(B * nd));
    ZKL_TRY(wait_stream());
    ZKL_TRY(comm_call(ctx, sg.comm->all_gather(sg.comm->user, cur_ev, ev_last.p, Bp * nd * sizeof(fe)), "all_gather(last FRI layer)"));
    cur_ev = ev_last.p; ev_sharded = false;
  }
  mark("fri layers");
  size_t rem_len = nd;  // (B*nd)/B coefficients
  ZKL_CHECK(ctx, rem_rev.alloc(rem_len));
  LAUNCH(ctx, k_fri_remainder, 1, 32 > (int)nd ? 32 : (int)nd, 0, cur_ev, (int)nd, (int)B, tN->wpow, N, fe_inv(fe_from_u64(3)),
         fe_inv(fe_from_u64(nd * B)), rem_rev.p);
  dev_buf<fe> rem_wide;
  const fe *rem_w = rem_rev.p;
  if (E == 2) {
    ZKL_CHECK(ctx, rem_wide.alloc(2 * rem_len));
    LAUNCH(ctx, k_spread_columns, ceil_div(2 * rem_len, 256), 256, 0, rem_rev.p, (size_t)1, rem_len, rem_wide.p);
    rem_w = rem_wide.p;
  }
  if (quad) LAUNCH(ctx, k_quad_hash_elements_single, 1, ZKL_QUAD_GROUP, 0, rem_w, (int)(rem_len * E), fri_roots.p + nl);
  else LAUNCH(ctx, k_hash_elements_single, 1, ZKL_COOP_GROUP, 0, rem_w, (int)(rem_len * E), fri_roots.p + nl);
  ZKL_TRY(reseed_draw(fri_roots.p + nl, nullptr, 0, nullptr));
  STAGE_MARK();

  // ---- 10. grinding (smallest nonce >= 1) and 11. query positions: enqueued back to back, ONE host wait for both.
  // The search window is 16x the expected nonce (miss probability e^-16): the positions are drawn right behind it from the
  // nonce on the device; the host only looks at the nonce after the wait and repeats with the next window on a miss.
  dev_buf<unsigned long long> nonce;
  ZKL_CHECK(ctx, nonce.alloc(1));
  unsigned long long nonce_h = 0;
  if (o->grinding > 40) { ctx->set_error("grinding factor too large"); return 1; }
  dev_buf<uint32_t> pos_dev, pos_cnt;
  ZKL_CHECK(ctx, pos_dev.alloc(256)); ZKL_CHECK(ctx, pos_cnt.alloc(1));
  dev_buf<fe> qseed, qdraw;
  ZKL_CHECK(ctx, qseed.alloc(1)); ZKL_CHECK(ctx, qdraw.alloc(o->num_queries));
  struct mailbox { unsigned long long nonce; uint32_t cnt, pad; uint32_t pos[256]; fe troot, croot, z, flag; };
  mailbox *mb = (mailbox *)stage_alloc(ctx, sizeof(mailbox));
  fe *ood_h = (fe *)stage_alloc(ctx, 2 * (Wt + C) * sizeof(fe)), *roots_h = (fe *)stage_alloc(ctx, (nl + 1) * sizeof(fe));
  fe *rem_h = (fe *)stage_alloc(ctx, rem_len * sizeof(fe));
  if (!mb || !ood_h || !roots_h || !rem_h) { ctx->set_error("staging buffer exhausted"); return 1; }
  uint64_t base = 1;
  for (int attempt = 0;; attempt++) {
    if (o->grinding > 0) {
      if (attempt == 0) ZKL_CHECK(ctx, cudaMemsetAsync(nonce.p, 0xFF, 8, st));
      uint64_t target = std::min<uint64_t>((uint64_t)16 << o->grinding, (uint64_t)1 << 30);
      if (ctx->hash_path == 0) {
        // one persistent launch; CTAs stop at the first tile above the best nonce found so far
        LAUNCH_TC_TILES(ctx, k_tc_grind, target, coin.p, base, (uint32_t)target, o->grinding, nonce.p);
        base += target;
      } else {
        const uint32_t batch = 1u << 15;
        int nb = (int)std::min<uint64_t>(std::max<uint64_t>(target / batch, 1), 512);
        for (int b = 0; b < nb; b++, base += batch)
          LAUNCH(ctx, k_grind, ceil_div((size_t)batch * ZKL_COOP_GROUP, ZKL_COOP_THREADS), ZKL_COOP_THREADS, 0, coin.p, base, batch, o->grinding, nonce.p);
      }
    } else {
      ZKL_CHECK(ctx, cudaMemsetAsync(nonce.p, 0, 8, st));
    }
    if (attempt == 0) STAGE_MARK();
    // seed' = merge_with_int(seed, nonce); the draws in parallel; mask + sort + de-duplicate
    if (quad) {
      LAUNCH(ctx, k_seed_with_nonce<ZKL_QUAD_GROUP>, 1, ZKL_QUAD_GROUP, 0, coin.p, nonce.p, qseed.p);
      LAUNCH_QUAD(ctx, k_quad_coin_draw, o->num_queries, qseed.p, (uint64_t)0, (int)o->num_queries, qdraw.p);
    } else {
      LAUNCH(ctx, k_seed_with_nonce<ZKL_COOP_GROUP>, 1, ZKL_COOP_GROUP, 0, coin.p, nonce.p, qseed.p);
      LAUNCH_COOP(ctx, k_coin_draw, o->num_queries, qseed.p, (uint64_t)0, (int)o->num_queries, qdraw.p);
    }
    LAUNCH(ctx, k_positions_finish, 1, 64, 0, qdraw.p, (int)o->num_queries, (uint32_t)(N - 1), pos_dev.p, pos_cnt.p);
    ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->nonce, nonce.p, 8, cudaMemcpyDeviceToHost, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(mb->pos, pos_dev.p, sizeof mb->pos, cudaMemcpyDeviceToHost, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->cnt, pos_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    if (attempt == 0) {
      ZKL_CHECK(ctx, cudaMemcpyAsync(ood_h, ood_w, 2 * (Wt + C) * sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(roots_h, fri_roots.p, (nl + 1) * sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(rem_h, rem_rev.p, rem_len * sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->troot, tl.m.nodes.p + 1, sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->croot, cmt.m.nodes.p + 1, sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->z, z.p, sizeof(fe), cudaMemcpyDeviceToHost, st));
      ZKL_CHECK(ctx, cudaMemcpyAsync(&mb->flag, flag.p, sizeof(fe), cudaMemcpyDeviceToHost, st));
    }
    ZKL_TRY(ctx_wait(ctx));
    nonce_h = mb->nonce;
    if (o->grinding == 0 || nonce_h != ~0ull) break;
  }
  const uint32_t *pos_h = mb->pos; const uint32_t cnt_h = mb->cnt;
  const fe troot = mb->troot, croot = mb->croot, z_h = mb->z, flag_h = mb->flag;
  std::vector<uint32_t> pos(pos_h, pos_h + cnt_h);

  // One gather for everything the proof opens: queried rows, FRI layer values and all Merkle nodes.
  // Phase 1 lists (source buffer, element index) in the exact order the serializer consumes them.
  const fe **g_src = (const fe **)stage_alloc(ctx, max_open * sizeof(fe *));
  uint32_t *g_idx = (uint32_t *)stage_alloc(ctx, max_open * 4);
  fe *got = (fe *)stage_alloc(ctx, max_open * sizeof(fe));
  if (!g_src || !g_idx || !got) { ctx->set_error("staging buffer exhausted"); return 1; }
  size_t g_cnt = 0;
  bool g_overflow = false;
  dev_buf<fe> zero_fe;
  if (sh) { ZKL_CHECK(ctx, zero_fe.alloc(1)); ZKL_CHECK(ctx, cudaMemsetAsync(zero_fe.p, 0, sizeof(fe), st)); }
  auto want = [&](const fe *src, uint32_t idx) { if (g_cnt < max_open) { g_src[g_cnt] = src; g_idx[g_cnt] = idx; g_cnt++; } else g_overflow = true; };
  // column-sharded prover: every opened element is contributed by exactly ONE rank and the others put a zero — its owner for sharded
  // data, rank 0 for what every rank holds — so that a single sum over the ranks assembles the openings (exchange 3c)
  auto want_rep = [&](const fe *src, uint32_t idx) { if (!sh || rk == 0) want(src, idx); else want(zero_fe.p, 0); };
  struct tree_job { std::vector<std::vector<node_ref>> plan; int depth; };
  auto plan_matrix = [&](const committed_matrix &m, tree_job &tj) {
    const bool sharded_cols = sh && &m == &tl.m;   // the trace matrix holds this rank's columns only: the others are opened as zero and summed in
    const bool sharded_cosets = sh && &m == &cmt.m;   // the composition matrix holds this rank's cosets only, compact
    for (uint32_t p : pos) {
      uint32_t s = (uint32_t)((p % m.blowup) * m.n + p / m.blowup);
      if (sharded_cosets) {
        const size_t kc = p % m.blowup;
        const bool mine = kc >= k0 && kc < k0 + Bp;
        for (size_t c = 0; c < m.width; c++) { if (mine) want(m.lde.p + c * (Bp * m.n), (uint32_t)((kc - k0) * m.n + p / m.blowup)); else want(zero_fe.p, 0); }
      }
      else if (!sharded_cols) for (size_t c = 0; c < m.width; c++) want_rep(m.lde.p + c * m.N, s);   // per-column base: width * N exceeds 32 bits for large segments
      else for (size_t c = 0; c < W; c++) { if (c >= col0 && c < col0 + Wl) want(m.lde.p + (c - col0) * m.N, s); else want(zero_fe.p, 0); }
    }
    tj.plan = batch_proof_plan(pos, m.N); tj.depth = ilog2_sz(m.N);
    for (auto &l : tj.plan) for (auto &r : l) want_rep(r.leaf ? m.leaves.p : m.nodes.p, r.index);
  };
  tree_job tj_trace, tj_cons; std::vector<tree_job> tj_fri(nl);
  std::vector<size_t> fri_counts(nl);
  plan_matrix(tl.m, tj_trace);
  plan_matrix(cmt.m, tj_cons);
  {
    std::vector<uint32_t> fp = pos; size_t dom = N; size_t ndl = n;
    for (size_t d = 0; d < nl; d++) {
      fp = fold_positions_host(fp, dom);
      size_t half = dom / 2, half_rows = ndl / 2;
      const fe *ev = layer_ev[d];
      for (uint32_t p : fp) {  // values ev[p], ev[p + half]; storage index of natural r: (r % B)*ndl + r / B
        if (!f_sharded[d]) {
          uint32_t s0 = (uint32_t)((p % B) * ndl + p / B);
          want_rep(ev, s0); want_rep(ev, s0 + (uint32_t)half_rows);
        } else if (p % B >= k0 && p % B < k0 + Bp) {   // the rank that owns the coset holds both values
          uint32_t s0 = (uint32_t)((p % B - k0) * ndl + p / B);
          want(ev, s0); want(ev, s0 + (uint32_t)half_rows);
        } else { want(zero_fe.p, 0); want(zero_fe.p, 0); }
      }
      fri_counts[d] = fp.size();
      tj_fri[d].plan = batch_proof_plan(fp, half); tj_fri[d].depth = ilog2_sz(half);
      for (auto &l : tj_fri[d].plan) for (auto &r : l) want_rep(r.leaf ? f_leaves[d].p : f_nodes[d].p, r.index);
      dom = half; ndl = half_rows;
    }
  }
  if (g_overflow) { ctx->set_error("opening plan larger than its bound"); return 1; }
  {
    dev_buf<const fe *> dsrc; dev_buf<uint32_t> didx; dev_buf<fe> dout;
    ZKL_CHECK(ctx, dsrc.alloc(g_cnt)); ZKL_CHECK(ctx, didx.alloc(g_cnt)); ZKL_CHECK(ctx, dout.alloc(g_cnt));
    ZKL_CHECK(ctx, cudaMemcpyAsync(dsrc.p, g_src, g_cnt * sizeof(fe *), cudaMemcpyHostToDevice, st));
    ZKL_CHECK(ctx, cudaMemcpyAsync(didx.p, g_idx, g_cnt * 4, cudaMemcpyHostToDevice, st));
    LAUNCH(ctx, k_gather_multi, ceil_div(g_cnt, 256), 256, 0, dsrc.p, didx.p, (int)g_cnt, dout.p);
    if (sh) {   // exchange (3c): one contributor per element, zeros elsewhere: a u64 sum assembles all openings exactly
      ZKL_TRY(wait_stream());
      ZKL_TRY(comm_call(ctx, sh->comm->all_reduce_sum_u64(sh->comm->user, dout.p, 2 * g_cnt), "all_reduce(openings)"));
    }
    ZKL_CHECK(ctx, cudaMemcpyAsync(got, dout.p, g_cnt * sizeof(fe), cudaMemcpyDeviceToHost, st));
    ZKL_TRY(ctx_wait(ctx));
  }
  size_t gp = 0;  // consumption cursor

  byte_writer wr;
  write_context_bytes(wr, W, n, *o, ncc);
  wr.u8((uint8_t)pos.size());
  wr.u16((uint16_t)(32 * (2 + nl + 1)));
  wr.digest(troot); wr.digest(croot);
  for (size_t i = 0; i <= nl; i++) wr.digest(roots_h[i]);
  auto take_lists = [&](const tree_job &tj) {
    std::vector<std::vector<fe>> lists;
    for (auto &l : tj.plan) { lists.emplace_back(got + gp, got + gp + l.size()); gp += l.size(); }
    return lists;
  };
  auto write_queries = [&](const committed_matrix &m, const tree_job &tj) {
    size_t cnt = pos.size() * (&m == &tl.m ? W : m.width);
    const fe *rows = got + gp; gp += cnt;
    byte_writer pw; write_batch_proof_bytes(pw, tj.depth, take_lists(tj));
    wr.usize(cnt * 16); wr.bytes(rows, cnt * 16);
    wr.usize(pw.b.size()); wr.bytes(pw.b.data(), pw.b.size());
  };
  write_queries(tl.m, tj_trace);
  write_queries(cmt.m, tj_cons);
  // OodFrame: trace states (current || next), constraint evaluations (current || next)
  wr.u16((uint16_t)(2 * Wt * 16));
  wr.bytes(ood_h, Wt * 16); wr.bytes(ood_h + Wt + C, Wt * 16);
  wr.u16((uint16_t)(2 * C * 16));
  wr.bytes(ood_h + Wt, C * 16); wr.bytes(ood_h + 2 * Wt + C, C * 16);
  // FriProof
  wr.u8((uint8_t)nl);
  for (size_t d = 0; d < nl; d++) {
    size_t cnt = 2 * fri_counts[d];
    const fe *vals = got + gp; gp += cnt;
    byte_writer pw; write_batch_proof_bytes(pw, tj_fri[d].depth, take_lists(tj_fri[d]));
    wr.u32((uint32_t)(cnt * 16 * E));
    if (E == 1) wr.bytes(vals, cnt * 16);
    else for (size_t i = 0; i < cnt; i++) wr.digest(vals[i]);    // an E element (a, 0): 16 bytes + 16 zero bytes
    wr.u32((uint32_t)pw.b.size()); wr.bytes(pw.b.data(), pw.b.size());
  }
  wr.u16((uint16_t)(rem_len * 16 * E));
  for (size_t i = 0; i < rem_len; i++) { if (E == 1) wr.felt(rem_h[i]); else wr.digest(rem_h[i]); }
  wr.u8(1);
  wr.u64(nonce_h);
  STAGE_MARK();
#undef STAGE_MARK

  if (summary) {
    memset(summary, 0, sizeof *summary);
    digest_from_fe_host(troot, summary->trace_root);
    digest_from_fe_host(croot, summary->constraint_root);
    summary->num_fri_roots = (uint32_t)(nl + 1);
    for (size_t i = 0; i <= nl; i++) digest_from_fe_host(roots_h[i], summary->fri_roots + 32 * i);
    memcpy(&summary->z, z_h.l, 16);
    summary->pow_nonce = nonce_h;
    summary->num_unique_queries = (uint32_t)pos.size();
    for (size_t i = 0; i < pos.size(); i++) summary->positions[i] = pos[i];
    ZKL_CHECK(ctx, cudaEventSynchronize(ctx->ev[evi - 1]));
    for (int i = 0; i + 1 < evi && i < 9; i++) cudaEventElapsedTime(&summary->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    cudaEventElapsedTime(&summary->stage_ms[9], ctx->ev[12], ctx->ev[13]);   // trace LDE kernel alone
    cudaEventElapsedTime(&summary->stage_ms[10], ctx->ev[13], ctx->ev[14]);  // trace row-hash kernel alone
    summary->constraints_violated = fe_is_zero(flag_h) ? 0 : 1;
  }
  if (out_len) {
    if (out_proof && wr.b.size() > *out_len) { *out_len = wr.b.size(); ctx->set_error("proof buffer too small"); return ZKL_ERR_BUFFER_TOO_SMALL; }
    if (out_proof) memcpy(out_proof, wr.b.data(), wr.b.size());
    *out_len = wr.b.size();
  }
  // composition degree overflow = some constraint does not vanish on the trace domain: the bytes above do not verify
  if (!fe_is_zero(flag_h)) { ctx->set_error("the trace does not satisfy the AIR (composition polynomial degree overflow): the proof will not verify"); return ZKL_ERR_CONSTRAINTS_VIOLATED; }
  return 0;
}

extern "C" int zkl_prove_segment_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n, const zkl_air_params *air,
                                     const zkl_fe *seed_elems, size_t num_seed, const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                                     zkl_proof_summary *summary) {
  return prove_segment_impl(ctx, (const fe *)trace_dev, width, n, air, seed_elems, num_seed, opts, out_proof, out_len, summary);
}
extern "C" int zkl_prove_segment(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n, const zkl_air_params *air, const zkl_fe *seed_elems,
                                 size_t num_seed, const zkl_options *opts, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary) {
  ZKL_TRY(enter(ctx));
  dev_buf<fe> tr; ZKL_CHECK(ctx, tr.alloc(width * n));
  ZKL_CHECK(ctx, cudaMemcpyAsync(tr.p, trace, width * n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  return prove_segment_impl(ctx, tr.p, width, n, air, seed_elems, num_seed, opts, out_proof, out_len, summary);
}

// the same with ZkLispAir::new on this side of the boundary (host_air.inl): nothing about a segment is cached between calls
static int prove_segment_pi_impl(zkl_ctx *ctx, const fe *trace_dev, size_t width, size_t n, const zkl_program *program, const zkl_public_inputs *pi,
                                 const zkl_options *opts, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary) {
  auto t0 = std::chrono::steady_clock::now();
  built_air b;
  if (build_air_host(program, pi, width, n, opts, b, false)) { ctx->set_error(b.err); return 1; }   // closed form: the device expands the assertion list
  const float host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  int rc = prove_segment_impl(ctx, trace_dev, width, n, &b.p, b.seed.data(), b.seed.size(), opts, out_proof, out_len, summary, nullptr, &b.sched);
  if (summary && (rc == 0 || rc == ZKL_ERR_CONSTRAINTS_VIOLATED)) summary->stage_ms[11] = host_ms;
  return rc;
}
extern "C" int zkl_prove_segment_pi_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n, const zkl_program *program,
                                        const zkl_public_inputs *pi, const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                                        zkl_proof_summary *summary) {
  return prove_segment_pi_impl(ctx, (const fe *)trace_dev, width, n, program, pi, opts, out_proof, out_len, summary);
}
extern "C" int zkl_prove_segment_pi(zkl_ctx *ctx, const zkl_fe *trace, size_t width, size_t n, const zkl_program *program,
                                    const zkl_public_inputs *pi, const zkl_options *opts, uint8_t *out_proof, size_t *out_len,
                                    zkl_proof_summary *summary) {
  ZKL_TRY(enter(ctx));
  dev_buf<fe> tr; ZKL_CHECK(ctx, tr.alloc(width * n));
  // the trace goes first so that its transfer overlaps the host-side AIR construction
  ZKL_CHECK(ctx, cudaMemcpyAsync(tr.p, trace, width * n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  return prove_segment_pi_impl(ctx, tr.p, width, n, program, pi, opts, out_proof, out_len, summary);
}

// ---- column-sharded proof of one segment (include/zkl_b200.h: zkl_comm) ----
extern "C" int zkl_shard_columns(const zkl_options *o, size_t width, int rank, int world, size_t *first_col, size_t *local_width) {
  if (world < 1 || rank < 0 || rank >= world || width == 0) return 1;
  const size_t ps = partition_size_of(o, width), P = (width + ps - 1) / ps;
  if (P < (size_t)world || (world > 1 && P < 2)) return 1;   // every rank owns at least one whole partition
  const size_t f = (size_t)rank * P / world, l = (size_t)(rank + 1) * P / world;
  const size_t lo = std::min(width, f * ps), hi = std::min(width, l * ps);
  if (hi <= lo) return 1;
  *first_col = lo; *local_width = hi - lo;
  return 0;
}
extern "C" int zkl_set_shard_min_leaves(zkl_ctx *ctx, size_t leaves) { ctx->shard_min_leaves = leaves; return 0; }
extern "C" int zkl_prove_segment_sharded_dev(zkl_ctx *ctx, const void *local_trace_dev, size_t first_col, size_t local_width, size_t width, size_t n,
                                             const zkl_air_params *air, const zkl_fe *seed_elems, size_t num_seed, const zkl_options *opts,
                                             const zkl_comm *comm, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary) {
  if (!comm) { ctx->set_error("zkl_comm is required"); return 1; }
  shard_view sh{comm, first_col, local_width};
  return prove_segment_impl(ctx, (const fe *)local_trace_dev, width, n, air, seed_elems, num_seed, opts, out_proof, out_len, summary, &sh);
}
extern "C" int zkl_prove_segment_sharded_pi_dev(zkl_ctx *ctx, const void *local_trace_dev, size_t first_col, size_t local_width, size_t width, size_t n,
                                                const zkl_program *program, const zkl_public_inputs *pi, const zkl_options *opts,
                                                const zkl_comm *comm, uint8_t *out_proof, size_t *out_len, zkl_proof_summary *summary) {
  if (!comm) { ctx->set_error("zkl_comm is required"); return 1; }
  auto t0 = std::chrono::steady_clock::now();
  built_air b;
  if (build_air_host(program, pi, width, n, opts, b, false)) { ctx->set_error(b.err); return 1; }
  const float host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  shard_view sh{comm, first_col, local_width};
  int rc = prove_segment_impl(ctx, (const fe *)local_trace_dev, width, n, &b.p, b.seed.data(), b.seed.size(), opts, out_proof, out_len, summary, &sh, &b.sched);
  if (summary && (rc == 0 || rc == ZKL_ERR_CONSTRAINTS_VIOLATED)) summary->stage_ms[11] = host_ms;
  return rc;
}
// world = 1: the "collectives" are device-to-device copies (all_gather / all_to_all of one block) and a no-op (sum over one rank)
static int loop_copy(void *, const void *send, void *recv, size_t bytes) {
  // a device-to-device cudaMemcpy returns before the copy has run: wait for the legacy stream it was put on
  return cudaMemcpy(recv, send, bytes, cudaMemcpyDeviceToDevice) == cudaSuccess && cudaStreamSynchronize(cudaStreamLegacy) == cudaSuccess ? 0 : 1;
}
static int loop_sum(void *, void *, size_t) { return 0; }
extern "C" const zkl_comm *zkl_comm_loopback(void) {
  static const zkl_comm c{0, 1, nullptr, loop_copy, loop_copy, loop_sum};
  return &c;
}

extern "C" int zkl_dev_alloc(zkl_ctx *ctx, size_t bytes, void **out_dev) { ZKL_CHECK(ctx, cudaSetDevice(ctx->device)); ZKL_CHECK(ctx, cudaMalloc(out_dev, bytes)); return 0; }
extern "C" int zkl_dev_free(zkl_ctx *ctx, void *dev) { ZKL_CHECK(ctx, cudaSetDevice(ctx->device)); ZKL_CHECK(ctx, cudaFree(dev)); return 0; }
extern "C" int zkl_dev_upload(zkl_ctx *ctx, void *dev, const void *host, size_t bytes) {
  ZKL_CHECK(ctx, cudaSetDevice(ctx->device));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int zkl_dev_download(zkl_ctx *ctx, void *host, const void *dev, size_t bytes) {
  ZKL_CHECK(ctx, cudaSetDevice(ctx->device));
  ZKL_CHECK(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// column-sharded trace commitment (SURVEY.md §8e "single huge segment"): building blocks on device pointers.
// A rank owns whole column partitions (agg/child.rs:1031-1044: leaf = merge_many(hash_elements(partition_k))), so
// iNTT, LDE and partition hashing are local; what crosses NVLink is one 16-byte digest per (row, partition).
// All calls are stream-ordered on the context's stream and return after enqueueing (zkl_sync to wait).
// ------------------------------------------------------------------------------------------------
extern "C" int zkl_lde_dev(zkl_ctx *ctx, const void *cols_dev, size_t ncols, size_t n, size_t blowup, void *out_polys_dev, void *out_lde_dev) {
  ZKL_TRY(enter(ctx));
  ZKL_TRY(dev_interpolate(ctx, (const fe *)cols_dev, (fe *)out_polys_dev, ncols, n));
  return dev_lde(ctx, (const fe *)out_polys_dev, (fe *)out_lde_dev, ncols, n, blowup);
}
extern "C" int zkl_hash_partitions_dev(zkl_ctx *ctx, const void *lde_dev, size_t ncols, size_t n, size_t blowup, size_t partition_size,
                                       void *out_digests_dev) {
  ZKL_TRY(need_suite(ctx));
  if (partition_size == 0 || ncols == 0) { ctx->set_error("empty partition"); return 1; }
  return dev_hash_partitions(ctx, (const fe *)lde_dev, ncols, n, blowup, partition_size, (fe *)out_digests_dev);
}
extern "C" int zkl_merge_partition_digests_dev(zkl_ctx *ctx, const void *digests_dev, size_t nparts, size_t row_stride, size_t first_row,
                                               size_t rows, void *out_leaves_dev) {
  ZKL_TRY(need_suite(ctx));
  if (nparts < 2) { ctx->set_error("merge_many leaves need at least two partitions (a single partition digest IS the leaf)"); return 1; }
  return dev_merge_partitions(ctx, (const fe *)digests_dev + first_row, row_stride, rows, nparts, (fe *)out_leaves_dev);
}
extern "C" int zkl_merkle_dev(zkl_ctx *ctx, const void *leaves_dev, size_t count, void *out_nodes_dev) {
  ZKL_TRY(need_suite(ctx));
  ZKL_CHECK(ctx, cudaMemsetAsync(out_nodes_dev, 0, sizeof(fe), ctx->stream));
  return dev_merkle(ctx, (const fe *)leaves_dev, (fe *)out_nodes_dev, count);
}
extern "C" size_t zkl_partition_size(const zkl_options *opts, size_t width) { return partition_size_of(opts, width); }

extern "C" int zkl_sync(zkl_ctx *ctx) { ZKL_CHECK(ctx, cudaSetDevice(ctx->device)); ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }

// ------------------------------------------------------------------------------------------------
// measurement helpers (bench.py): stream timers and an IMAD.WIDE peak probe
// ------------------------------------------------------------------------------------------------
extern "C" int zkl_timer_start(zkl_ctx *ctx) {
  ZKL_TRY(enter(ctx));
  while (ctx->ev.size() < 18) { cudaEvent_t e; ZKL_CHECK(ctx, cudaEventCreate(&e)); ctx->ev.push_back(e); }
  ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[16], ctx->stream));
  return 0;
}
extern "C" int zkl_timer_stop(zkl_ctx *ctx, float *ms) {
  ZKL_CHECK(ctx, cudaSetDevice(ctx->device));
  ZKL_CHECK(ctx, cudaEventRecord(ctx->ev[17], ctx->stream));
  ZKL_CHECK(ctx, cudaEventSynchronize(ctx->ev[17]));
  ZKL_CHECK(ctx, cudaEventElapsedTime(ms, ctx->ev[16], ctx->ev[17]));
  return 0;
}
// dependency-free 32x32+64 multiply-accumulate chains: the integer-pipe roofline denominator
__global__ void __launch_bounds__(256) k_imad_probe(uint32_t seed, int iters, unsigned long long *sink) {
  unsigned long long a[8];
  uint32_t x = seed + threadIdx.x, y = seed * 2654435761u + blockIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = i + x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (unsigned long long)x * (uint32_t)(y + i) + a[i];   // IMAD.WIDE.U32
    x ^= (uint32_t)a[0];
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
  if (s == 0x123456789ull) *sink = s;
}
// returns wide multiply-accumulates per second (whole GPU)
extern "C" int zkl_imad_peak(zkl_ctx *ctx, double *wide_macs_per_s) {
  ZKL_TRY(enter(ctx));
  dev_buf<unsigned long long> sink; ZKL_CHECK(ctx, sink.alloc(1));
  const int iters = 4096, blocks = ctx->num_sms * 8;
  float best = 1e30f, ms = 0;
  for (int rep = 0; rep < 4; rep++) {
    ZKL_TRY(zkl_timer_start(ctx));
    LAUNCH(ctx, k_imad_probe, blocks, 256, 0, 12345u + rep, iters, sink.p);
    ZKL_TRY(zkl_timer_stop(ctx, &ms));
    if (rep && ms < best) best = ms;
  }
  *wide_macs_per_s = (double)blocks * 256.0 * iters * 8.0 / (best * 1e-3);
  return 0;
}
// Saturating tcgen05.mma kind::i8 loop: the tensor-pipe roofline denominator of the sponge kernels, measured instead of derived.
// One 128-thread CTA per SM; one thread issues `iters` MMAs (M 128, N 256, K 32, u8 x u8 -> s32, operands in shared memory,
// K-major, no swizzle — the same operand path as tcperm.cuh) alternating between two TMEM accumulators, one commit at the end.
#define TCPK_N 256
#define TCPK_IDESC ((2u << 4) | ((uint32_t)(TCPK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
__global__ void __launch_bounds__(128) k_tc_i8_peak(int iters, unsigned int *sink) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char *gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_tile = base, b_tile = base + 128 * 32, ctl = b_tile + TCPK_N * 32;
  uint32_t *fill = reinterpret_cast<uint32_t *>(gen);
  for (int i = threadIdx.x; i < (128 * 32 + TCPK_N * 32) / 4; i += 128) fill[i] = (uint32_t)(i + 1) * 2654435761u + blockIdx.x;
  uint32_t *ctl_gen = reinterpret_cast<uint32_t *>(gen + 128 * 32 + TCPK_N * 32);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 16), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) { mbar_init(ctl, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl_gen[4];
  if (threadIdx.x == 0) {
    const uint64_t adesc = tc_smem_desc(a_tile, 128 * 16, 128), bdesc = tc_smem_desc(b_tile, TCPK_N * 16, 128);
#pragma unroll 4
    for (int i = 0; i < iters; i++) tc_mma_i8(tmem + 256 * (i & 1), adesc, bdesc, TCPK_IDESC, i > 1);
    tc_commit(ctl);
  }
  mbar_wait(ctl, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem + ((uint32_t)(32 * (threadIdx.x >> 5)) << 16), v);
  tmem_wait_ld(v);
  if (v[threadIdx.x & 15] == 0x7fffffffu) *sink = v[0];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}
// returns dense u8 x u8 -> s32 tensor operations per second (2 per multiply-accumulate, whole GPU)
extern "C" int zkl_tc_i8_peak(zkl_ctx *ctx, double *ops_per_s) {
  ZKL_TRY(enter(ctx));
  dev_buf<unsigned int> sink; ZKL_CHECK(ctx, sink.alloc(1));
  const int iters = 16384, blocks = ctx->num_sms;
  const size_t smem = 128 * 32 + TCPK_N * 32 + 256;
  float best = 1e30f, ms = 0;
  for (int rep = 0; rep < 4; rep++) {
    ZKL_TRY(zkl_timer_start(ctx));
    LAUNCH(ctx, k_tc_i8_peak, blocks, 128, smem, iters, sink.p);
    ZKL_TRY(zkl_timer_stop(ctx, &ms));
    if (rep && ms < best) best = ms;
  }
  *ops_per_s = (double)blocks * iters * 2.0 * 128.0 * TCPK_N * 32.0 / (best * 1e-3);
  return 0;
}
extern "C" int zkl_set_bulk_threshold(zkl_ctx *ctx, long sponges) { ctx->bulk_threshold = sponges; return 0; }
extern "C" int zkl_set_coop_path(zkl_ctx *ctx, int path) {
  if (path < 0 || path > 2) { ctx->set_error("coop path must be 0 (automatic), 1 (16-thread groups) or 2 (64-thread groups)"); return 1; }
  ctx->coop_path = path; return 0;
}
extern "C" int zkl_set_ntt_path(zkl_ctx *ctx, int path) {
  if (path < 0 || path > 2) { ctx->set_error("ntt path must be 0 (tensor-core radix-16 LDE), 1 (integer pipe) or 2 (tensor-core, separate coset shift)"); return 1; }
  ctx->ntt_path = path; return 0;
}
extern "C" int zkl_set_hash_path(zkl_ctx *ctx, int path) {
  if (path != 0 && path != 1) { ctx->set_error("hash path must be 0 (tensor-core MDS) or 1 (integer pipe)"); return 1; }
  if (path == 0 && !ctx->tc_ok) { ctx->set_error("the loaded MDS constants exceed the tensor-core accumulator bound (TC_ACC_MAX)"); return 1; }
  ctx->hash_path = path; return 0;
}

// permutation throughput probes on device-resident data: kind 0 = thread-per-state kernel, 1 = cooperative kernel,
// 2 = tensor-core MDS kernel, 3 = 64-thread groups (kernels_quad.cuh)
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_coop_permute_batch(fe *states, size_t count) {
  __shared__ fe s_xch[(ZKL_COOP_THREADS / ZKL_COOP_GROUP) * 24];
  size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  const int l = threadIdx.x & (ZKL_COOP_GROUP - 1);
  size_t gi = g < count ? g : count - 1;
  fe st = l < 12 ? ld_fe(states + 12 * gi + l) : fe_zero();
  coop_permute(st, l, s_xch + (threadIdx.x / ZKL_COOP_GROUP) * 24, 0xFFFFu << (threadIdx.x & 16));
  if (g < count && l < 12) st_fe(states + 12 * g + l, st);
}
__global__ void __launch_bounds__(ZKL_QUAD_THREADS) k_quad_permute_batch(fe *states, size_t count) {
  __shared__ fe s_xch[(ZKL_QUAD_THREADS / ZKL_QUAD_GROUP) * 24];
  const quad_ctx q = quad_setup(s_xch);
  size_t g = quad_index(), gi = g < count ? g : count - 1;
  fe st = ld_fe(states + 12 * gi + q.l);
  quad_permute(st, q);
  if (g < count && q.writer) st_fe(states + 12 * g + q.l, st);
}
extern "C" int zkl_bench_permute(zkl_ctx *ctx, int kind, size_t count, int iters, float *ms_per_iter) {
  ZKL_TRY(need_suite(ctx));
  dev_buf<fe> d; ZKL_CHECK(ctx, d.alloc(count * 12));
  ZKL_CHECK(ctx, cudaMemsetAsync(d.p, 0x11, count * 12 * sizeof(fe), ctx->stream));
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    ZKL_TRY(zkl_timer_start(ctx));
    for (int i = 0; i < iters; i++) {
      if (kind == 2) LAUNCH_TC(ctx, k_tc_permute_batch, count, d.p, count);
      else if (kind == 0) LAUNCH(ctx, k_permute_batch, ceil_div(count, ZKL_HASH_THREADS), ZKL_HASH_THREADS, 0, d.p, count);
      else if (kind == 3) LAUNCH_QUAD(ctx, k_quad_permute_batch, count, d.p, count);
      else LAUNCH(ctx, k_coop_permute_batch, ceil_div(count * ZKL_COOP_GROUP, ZKL_COOP_THREADS), ZKL_COOP_THREADS, 0, d.p, count);
    }
    ZKL_TRY(zkl_timer_stop(ctx, &ms));
  }
  *ms_per_iter = ms / iters;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// segment driver helpers (callers of the hot path): the reference slices the full trace
// per segment on the host (vm/trace/mod.rs:367-384) and scans every row for gadget usage (prove.rs:1289-1392).
// With the full trace resident in HBM both become one kernel each.
// ------------------------------------------------------------------------------------------------
// out[c][row] = full[seg_to_full[c]][r_start + row]
__global__ void k_slice_segment(const fe *__restrict__ full, size_t full_rows, const uint32_t *__restrict__ seg_to_full, int seg_width,
                                size_t r_start, size_t seg_rows, fe *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)seg_width * seg_rows) return;
  size_t c = t / seg_rows, row = t - c * seg_rows;
  st_fe(out + t, ld_fe(full + (size_t)seg_to_full[c] * full_rows + r_start + row));
}
extern "C" int zkl_slice_segment_dev(zkl_ctx *ctx, const void *full_dev, size_t full_rows, const uint32_t *seg_to_full, size_t seg_width,
                                     size_t r_start, size_t seg_rows, void *out_dev) {
  ZKL_TRY(enter(ctx));
  if (r_start + seg_rows > full_rows) { ctx->set_error("segment out of bounds for the full trace"); return 1; }
  dev_buf<uint32_t> map; ZKL_CHECK(ctx, map.alloc(seg_width));
  ZKL_CHECK(ctx, cudaMemcpyAsync(map.p, seg_to_full, seg_width * 4, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_slice_segment, ceil_div(seg_width * seg_rows, 256), 256, 0, (const fe *)full_dev, full_rows, map.p, (int)seg_width, r_start,
         seg_rows, (fe *)out_dev);
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // `seg_to_full` is caller memory
  return 0;
}
// Poseidon gadget rows of the execution trace, batched over levels (vm/trace/poseidon.rs:24-96 `apply_level_absorb`): the trace
// builder's only dense arithmetic (27 rounds x 168 field products per active level). One thread per level; row 32L is the map
// row (inputs in lanes 0..9, the program suite's domain tags in lanes 10, 11), row 32L+1+j the state BEFORE round j, row 32L+28
// (final) and the pad rows 29..31 the state after round 26. The suite is the PROGRAM's (suite_id = program_id), not the hasher's.
struct pose_suite_dev { fe dom[2]; fe mds[12][12]; fe rc[ZKL_ROUNDS][12]; };
__global__ void __launch_bounds__(64) k_poseidon_levels(fe *__restrict__ trace, size_t n_rows, uint32_t lanes_start, const pose_suite_dev *__restrict__ s,
                                                        const uint32_t *__restrict__ levels, const fe *__restrict__ inputs, int count) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const size_t base = (size_t)levels[t] * 32;
  fe st[12];
#pragma unroll
  for (int i = 0; i < 10; i++) st[i] = ld_fe(inputs + (size_t)t * 10 + i);
  st[10] = ld_fe(&s->dom[0]); st[11] = ld_fe(&s->dom[1]);
  fe *lane0 = trace + (size_t)lanes_start * n_rows + base;
#define PUT_ROW(r) for (int i_ = 0; i_ < 12; i_++) st_fe(lane0 + (size_t)i_ * n_rows + (r), st[i_])
  PUT_ROW(0);
#pragma unroll 1
  for (int j = 0; j < ZKL_ROUNDS; j++) {
    PUT_ROW(1 + j);
    fe c3[12];
#pragma unroll
    for (int i = 0; i < 12; i++) c3[i] = fe_cube(st[i]);
#pragma unroll 1
    for (int i = 0; i < 12; i++) {
      wide_acc w; wacc_zero(w);
#pragma unroll
      for (int k = 0; k < 12; k++) wacc_mad(w, c3[k], ld_fe(&s->mds[i][k]));
      st[i] = fe_add(wacc_reduce(w), ld_fe(&s->rc[j][i]));
    }
  }
  for (int r = 28; r < 32; r++) { PUT_ROW(r); }
#undef PUT_ROW
}
extern "C" int zkl_poseidon_levels_dev(zkl_ctx *ctx, void *trace_dev, size_t n_rows, uint32_t lanes_start, const zkl_suite *suite,
                                       const uint32_t *levels, const zkl_fe *inputs, size_t count) {
  ZKL_TRY(enter(ctx));
  if (count == 0) return 0;
  if (count > (1u << 30)) { ctx->set_error("too many levels"); return 1; }
  for (size_t i = 0; i < count; i++)
    if (((size_t)levels[i] + 1) * 32 > n_rows) { ctx->set_error("level outside the trace"); return 1; }
  pose_suite_dev h;
  for (int i = 0; i < 2; i++) h.dom[i] = fe_from_abi(suite->dom[i]);
  for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) h.mds[i][j] = fe_from_abi(suite->mds[i][j]);
  for (int r = 0; r < ZKL_ROUNDS; r++) for (int j = 0; j < 12; j++) h.rc[r][j] = fe_from_abi(suite->rc[r][j]);
  dev_buf<pose_suite_dev> ds; dev_buf<uint32_t> dl; dev_buf<fe> di;
  ZKL_CHECK(ctx, ds.alloc(1)); ZKL_CHECK(ctx, dl.alloc(count)); ZKL_CHECK(ctx, di.alloc(count * 10));
  ZKL_CHECK(ctx, cudaMemcpyAsync(ds.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dl.p, levels, count * 4, cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(di.p, inputs, count * 10 * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_poseidon_levels, ceil_div(count, (size_t)64), 64, 0, (fe *)trace_dev, n_rows, lanes_start, ds.p, dl.p, di.p, (int)count);
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // suite, levels and inputs are caller / stack memory
  return 0;
}
// compute_vm_usage_mask_for_trace (prove.rs:1289-1392): out[0] = vm usage mask, out[1] = ram delta_clk bit mask
struct usage_cols {
  uint32_t op_assert, op_select, op_assert_bit, op_assert_range, op_divmod, op_mulwide, op_div128, op_eq, op_sponge;
  uint32_t ram_sorted, ram_s_addr, gadget_b_start, has_ram;
};
__global__ void k_usage_mask(const fe *__restrict__ tr, size_t n, usage_cols c, unsigned int *out) {
  size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int mask = 0, bits = 0;
  if (row < n) {
    auto nz = [&](uint32_t col, size_t r) { return !fe_is_zero(ld_fe(tr + (size_t)col * n + r)); };
    bool at_final = (row % 32) == 28;
    if (at_final && (nz(c.op_assert, row) || nz(c.op_select, row))) mask |= 1u << 0;
    if (at_final && nz(c.op_assert_bit, row)) mask |= 1u << 1;
    if (at_final && nz(c.op_assert_range, row)) mask |= 1u << 2;
    if (at_final && nz(c.op_divmod, row)) mask |= 1u << 3;
    if (at_final && nz(c.op_mulwide, row)) mask |= 1u << 4;
    if (at_final && nz(c.op_div128, row)) mask |= 1u << 5;
    if (at_final && nz(c.op_eq, row)) mask |= 1u << 6;
    if (nz(c.op_sponge, row)) mask |= 1u << 7;
    if (c.has_ram && row + 1 < n && nz(c.ram_sorted, row) && nz(c.ram_sorted, row + 1) &&
        fe_eq(ld_fe(tr + (size_t)c.ram_s_addr * n + row), ld_fe(tr + (size_t)c.ram_s_addr * n + row + 1))) {
      mask |= 1u << 8;
      for (int i = 0; i < 32; i++) if (nz(c.gadget_b_start + i, row)) bits |= 1u << i;
    }
  }
  mask = __reduce_or_sync(0xffffffffu, mask); bits = __reduce_or_sync(0xffffffffu, bits);
  if ((threadIdx.x & 31) == 0) { if (mask) atomicOr(out, mask); if (bits) atomicOr(out + 1, bits); }
}
extern "C" int zkl_usage_mask_dev(zkl_ctx *ctx, const void *trace_dev, size_t width, size_t n, const zkl_columns *cols, int has_ram,
                                  uint32_t *out_mask, uint32_t *out_ram_bits) {
  ZKL_TRY(enter(ctx));
  (void)width;
  usage_cols c;
  const uint32_t o = cols->op_start;
  c.op_eq = o + 6; c.op_select = o + 7; c.op_sponge = o + 8; c.op_assert = o + 9; c.op_assert_bit = o + 10; c.op_assert_range = o + 11;
  c.op_divmod = o + 12; c.op_div128 = o + 13; c.op_mulwide = o + 14;
  c.ram_sorted = cols->ram_start; c.ram_s_addr = cols->ram_start + 1; c.gadget_b_start = cols->gadget_b_start; c.has_ram = has_ram ? 1 : 0;
  dev_buf<unsigned int> d; ZKL_CHECK(ctx, d.alloc(2));
  ZKL_CHECK(ctx, cudaMemsetAsync(d.p, 0, 8, ctx->stream));
  LAUNCH(ctx, k_usage_mask, ceil_div(n, 256), 256, 0, (const fe *)trace_dev, n, c, d.p);
  unsigned int h[2];
  ZKL_CHECK(ctx, cudaMemcpyAsync(h, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  *out_mask = h[0]; *out_ram_bits = h[1];
  return 0;
}
// out[i] = trace[cols[i]][rows[i]] (boundary values: prove.rs:1197-1287 compute_segment_boundary_bytes)
__global__ void k_read_cells(const fe *__restrict__ tr, size_t n, const uint32_t *__restrict__ cols, const uint32_t *__restrict__ rows, int k,
                             fe *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < k) st_fe(out + t, ld_fe(tr + (size_t)cols[t] * n + rows[t]));
}
extern "C" int zkl_read_cells_dev(zkl_ctx *ctx, const void *trace_dev, size_t n, const uint32_t *cols, const uint32_t *rows, size_t k,
                                  zkl_fe *out) {
  ZKL_TRY(enter(ctx));
  dev_buf<uint32_t> dc, dr; dev_buf<fe> o;
  ZKL_CHECK(ctx, dc.alloc(k)); ZKL_CHECK(ctx, dr.alloc(k)); ZKL_CHECK(ctx, o.alloc(k));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dc.p, cols, k * 4, cudaMemcpyHostToDevice, ctx->stream));
  ZKL_CHECK(ctx, cudaMemcpyAsync(dr.p, rows, k * 4, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(ctx, k_read_cells, ceil_div(k, 128), 128, 0, (const fe *)trace_dev, n, dc.p, dr.p, (int)k, o.p);
  ZKL_CHECK(ctx, cudaMemcpyAsync(out, o.p, k * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ZKL_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
