// kernels_stark.cuh — composition split, OOD evaluation, DEEP composition, FRI folding, the Fiat-Shamir
// coin, grinding and query extraction: stages 6-11 of winterfell::Prover::prove.
// In-repo restatements followed: transcript zk-lisp-proof-winterfell/src/agg/fs.rs:67-228; DEEP
// src/agg/trace.rs:1134-1193; FRI fold/remainder src/agg/trace.rs:752-950; positions src/agg/child.rs:1072-1099.
#pragma once
#include "kernels_ntt.cuh"

// ---------------------------------------------------------------- gathers
// rows of a coset-major matrix at natural positions: out[q][c] = m[c][storage(pos[q])]
__global__ void k_gather_rows(const fe *__restrict__ m, size_t N, size_t n, int blowup, int W, const uint32_t *__restrict__ pos,
                              int k, fe *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k * W) return;
  int q = t / W, c = t % W;
  size_t r = pos[q];
  size_t s = (r % blowup) * n + (r / blowup);
  st_fe(out + (size_t)q * W + c, ld_fe(m + (size_t)c * N + s));
}
// out[i] = src[idx[i]]
__global__ void k_gather_fe(const fe *__restrict__ src, const uint32_t *__restrict__ idx, int count, fe *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  st_fe(out + t, ld_fe(src + idx[t]));
}

// out[i] = src[i][idx[i]] — one launch gathers every opened row, FRI value and Merkle node of a proof
__global__ void k_gather_multi(const fe *const *__restrict__ src, const uint32_t *__restrict__ idx, int count, fe *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  st_fe(out + t, ld_fe(src[t] + idx[t]));
}

// ---------------------------------------------------------------- Fiat-Shamir coin (winter-crypto DefaultRandomCoin [WF-mem])
// The seed lives on the device; the draw counter is tracked by the host (it is data independent).
__global__ void k_coin_init(const fe *__restrict__ elems, int count, fe *seed) {  // seed = hash_elements(elems)
  fe d = coop_hash_elements(elems, 1, count);
  if (threadIdx.x == 0) st_fe(seed, d);
}
// out[i] = merge_with_int(seed, counter_base + 1 + i) as a field element (draw::<BaseElement>), i < count
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_coin_draw(const fe *__restrict__ seed, uint64_t counter_base, int count, fe *__restrict__ out) {
  int g = (blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  int gi = g < count ? g : count - 1;
  fe d = coop_merge_with_int(ld_fe(seed), counter_base + 1 + (uint64_t)gi);
  if (g < count && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) st_fe(out + g, d);
}
// hash_elements of `count` contiguous elements with one group (OOD frame digest, remainder commitment)
__global__ void k_hash_elements_single(const fe *__restrict__ elems, int count, fe *out) {
  fe d = coop_hash_elements(elems, 1, count);
  if (threadIdx.x == 0) st_fe(out, d);
}

// grinding: smallest nonce >= 1 with trailing_zeros(u64_le(merge_with_int(seed, nonce)[0..8])) >= bits
// (agg/fs.rs:202-207). One group per nonce; result = atomicMin over qualifying nonces. A batch exits at once
// when an earlier batch already found a smaller nonce.
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_grind(const fe *__restrict__ seed, uint64_t nonce_base, uint32_t count, uint32_t bits,
                                                            unsigned long long *result) {
  if (*(volatile unsigned long long *)result < nonce_base) return;
  uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  uint64_t nonce = nonce_base + (g < count ? g : count - 1);
  fe d = coop_merge_with_int(ld_fe(seed), nonce);
  if (g < count && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) {
    uint64_t head = ((uint64_t)d.l[1] << 32) | d.l[0];
    uint32_t tz = head == 0 ? 64 : (uint32_t)__ffsll((long long)head) - 1;
    if (tz >= bits) atomicMin(result, (unsigned long long)nonce);
  }
}
// ---------------------------------------------------------------- composition polynomial (CompositionPoly::new [WF-mem])
// After the per-coset inverse NTT, a[j1][m2] = w_M^(j1*m2) * sum_m1 w_ce^(j1*m1) * (h_{m1*n+m2} g^{m1*n+m2}).
// Thread m2 undoes the twiddle, runs the size-ce inverse DFT over j1 and removes the offset:
//   out[m1][m2] = h_{m1*n + m2}  for m1 < num_cols.  ginv_pow[j] = g^-j (j < n), ginv_n = g^-n.
__global__ void k_composition_finish(const fe *__restrict__ a, int n, int ce, int num_cols, const fe *__restrict__ wpowM, size_t M,
                                     const fe *__restrict__ ginv_pow, fe ginv_n, fe ce_inv, fe *__restrict__ out, fe *__restrict__ high_check) {
  int m2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (m2 >= n) return;
  fe v[16];
  for (int j1 = 0; j1 < ce; j1++) {
    size_t e = ((size_t)j1 * m2) % M;
    fe tw = ld_fe(wpowM + (e ? M - e : 0));
    v[j1] = fe_mul(ld_fe(a + (size_t)j1 * n + m2), tw);
  }
  fe scale = fe_mul(ld_fe(ginv_pow + m2), ce_inv);
  for (int m1 = 0; m1 < ce; m1++) {
    fe acc = fe_zero();
    for (int j1 = 0; j1 < ce; j1++) {
      size_t e = ((size_t)j1 * m1 * (M / ce)) % M;  // w_ce^(j1*m1) = w_M^(j1*m1*n)
      acc = fe_add(acc, fe_mul(v[j1], ld_fe(wpowM + (e ? M - e : 0))));
    }
    fe h = fe_mul(acc, scale);
    if (m1 < num_cols) st_fe(out + (size_t)m1 * n + m2, h);
    else if (!fe_is_zero(h)) high_check[0] = fe_one();  // degree overflow marker (invalid trace)
    scale = fe_mul(scale, ginv_n);
  }
}

// ---------------------------------------------------------------- OOD evaluation
// out[j] = (base * mulc)^j, j < n ; base read from device memory (z lives on the device)
__global__ void k_powers_dev(const fe *__restrict__ base, fe mulc, fe *__restrict__ out, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  st_fe(out + j, fe_pow(fe_mul(ld_fe(base), mulc), (uint64_t)j));
}
// Two out-of-domain points at once: part[(p * chunks + ch) * 2 + q] = sum over chunk ch of polys[p][j] * xpow[q * n + j], q = 0 (z), 1 (z g).
// grid = (chunks, polynomials): every coefficient is read ONCE for both points, and a long column is spread over many SMs (a
// 2^20-row column on one CTA is 16 MB through one SM). chunks == 1 writes the results straight to out0[p] / out1[p].
#define POLY_EVAL_CHUNK 16384
__global__ void __launch_bounds__(256) k_poly_eval2(const fe *__restrict__ polys, int n, const fe *__restrict__ xpow, fe *__restrict__ part,
                                                    fe *__restrict__ out0, fe *__restrict__ out1) {
  __shared__ fe s_part[2][256];
  const int chunks = gridDim.x, ch = blockIdx.x;
  const fe *poly = polys + (size_t)blockIdx.y * n;
  const int j_lo = ch * POLY_EVAL_CHUNK, j_hi = min(n, j_lo + POLY_EVAL_CHUNK);
  fe acc0 = fe_zero(), acc1 = fe_zero();
  for (int j0 = j_lo + threadIdx.x; j0 < j_hi; j0 += blockDim.x * 8) {
    wide_acc w0, w1; wacc_zero(w0); wacc_zero(w1);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int j = j0 + u * blockDim.x;
      if (j < j_hi) { fe c = ld_fe(poly + j); wacc_mad(w0, c, ld_fe(xpow + j)); wacc_mad(w1, c, ld_fe(xpow + n + j)); }
    }
    acc0 = fe_add(acc0, wacc_reduce(w0)); acc1 = fe_add(acc1, wacc_reduce(w1));
  }
  s_part[0][threadIdx.x] = acc0; s_part[1][threadIdx.x] = acc1;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_part[0][threadIdx.x] = fe_add(s_part[0][threadIdx.x], s_part[0][threadIdx.x + s]);
      s_part[1][threadIdx.x] = fe_add(s_part[1][threadIdx.x], s_part[1][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (chunks == 1) { st_fe(out0 + blockIdx.y, s_part[0][0]); st_fe(out1 + blockIdx.y, s_part[1][0]); }
    else { st_fe(part + ((size_t)blockIdx.y * chunks + ch) * 2, s_part[0][0]); st_fe(part + ((size_t)blockIdx.y * chunks + ch) * 2 + 1, s_part[1][0]); }
  }
}
__global__ void k_poly_eval2_finish(const fe *__restrict__ part, int chunks, int npolys, fe *__restrict__ out0, fe *__restrict__ out1) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * npolys) return;
  const int p = t >> 1, q = t & 1;
  fe a = fe_zero();
  for (int ch = 0; ch < chunks; ch++) a = fe_add(a, ld_fe(part + ((size_t)p * chunks + ch) * 2 + q));
  st_fe((q ? out1 : out0) + p, a);
}

// ---------------------------------------------------------------- DEEP composition (agg/trace.rs:1158-1193)
// frame layout (OOD evaluations, the order hashed for the reseed, agg/fs.rs:152-164):
//   ood[0..W) = T(z), ood[W..W+C) = H(z), ood[W+C..2W+C) = T(zg), ood[2W+C..2W+2C) = H(zg)
// az[0] = sum_i dc_i * (value at z), az[1] = same at z*g   (dc: W trace coefficients then C constraint ones)
__global__ void k_deep_consts(const fe *__restrict__ ood, const fe *__restrict__ dc, int W, int C, fe *az) {
  if (threadIdx.x || blockIdx.x) return;
  fe a = fe_zero(), b = fe_zero();
  for (int i = 0; i < W; i++) { a = fe_add(a, fe_mul(ld_fe(dc + i), ld_fe(ood + i))); b = fe_add(b, fe_mul(ld_fe(dc + i), ld_fe(ood + W + C + i))); }
  for (int j = 0; j < C; j++) { a = fe_add(a, fe_mul(ld_fe(dc + W + j), ld_fe(ood + W + j))); b = fe_add(b, fe_mul(ld_fe(dc + W + j), ld_fe(ood + 2 * W + C + j))); }
  st_fe(az, a); st_fe(az + 1, b);
}
// deep[s] = (A(x) - A_z)/(x - z) + (A(x) - A_zg)/(x - z g), A(x) = sum_i dc_i T_i(x) + sum_j dc_{W+j} H_j(x), coset-major
// inv[s] = 1 / ((x_s - z)(x_s - z g_n)) for every LDE point (storage order s), by Montgomery batch inversion: a CTA of 256
// threads takes 2048 points (8 consecutive per thread), multiplies them up (thread-local prefix products, then an exclusive
// prefix AND suffix scan of the 256 thread totals in shared memory), inverts the ONE block total (warp 0; a field inversion
// is 253 dependent products) and walks back down: ~7 products per point instead of 253.
#define DEEP_INV_PER_THREAD 8
__global__ void __launch_bounds__(256) k_deep_inverses(size_t N, int n, int blowup, const fe *__restrict__ z, fe g_n, const fe *__restrict__ wpowN,
                                                       fe *__restrict__ inv) {
  __shared__ fe s_pre[256], s_suf[256], s_tot;
  const int tid = threadIdx.x;
  const size_t s0 = ((size_t)blockIdx.x * 256 + tid) * DEEP_INV_PER_THREAD;
  const fe zz = ld_fe(z), zg = fe_mul(zz, g_n);
  fe v[DEEP_INV_PER_THREAD], p[DEEP_INV_PER_THREAD];
#pragma unroll
  for (int j = 0; j < DEEP_INV_PER_THREAD; j++) {
    size_t s = s0 + j;
    if (s < N) {
      size_t r = (s / n) + (size_t)blowup * (s % n);
      fe x = fe_mul(fe_from_u64(3), ld_fe(wpowN + r));
      v[j] = fe_mul(fe_sub(x, zz), fe_sub(x, zg));
    } else v[j] = fe_one();
    p[j] = j ? fe_mul(p[j - 1], v[j]) : v[j];
  }
  // inclusive scans of the thread totals: prefix in s_pre, suffix in s_suf (Hillis-Steele, 8 steps)
  s_pre[tid] = p[DEEP_INV_PER_THREAD - 1]; s_suf[tid] = p[DEEP_INV_PER_THREAD - 1];
  __syncthreads();
#pragma unroll 1
  for (int d = 1; d < 256; d <<= 1) {
    fe a = s_pre[tid], b = s_suf[tid];
    if (tid >= d) a = fe_mul(a, s_pre[tid - d]);
    if (tid + d < 256) b = fe_mul(b, s_suf[tid + d]);
    __syncthreads();
    s_pre[tid] = a; s_suf[tid] = b;
    __syncthreads();
  }
  if (tid < 32) {   // one warp inverts the block total (all its lanes run the same chain)
    fe t = fe_inv(s_pre[255]);
    if (tid == 0) s_tot = t;
  }
  __syncthreads();
  // 1 / (this thread's total) = (1 / block total) * (product of the totals before) * (product of the totals after)
  fe r = s_tot;
  if (tid > 0) r = fe_mul(r, s_pre[tid - 1]);
  if (tid < 255) r = fe_mul(r, s_suf[tid + 1]);
#pragma unroll
  for (int j = DEEP_INV_PER_THREAD - 1; j >= 0; j--) {
    fe iv = j ? fe_mul(r, p[j - 1]) : r;
    if (j) r = fe_mul(r, v[j]);
    if (s0 + j < N) st_fe(inv + s0 + j, iv);
  }
}
// Covers storage indices s_base .. s_base + count - 1 and writes deep[s - s_base] (the whole domain: s_base = 0, count = N); with
// `accumulate` the value is added to what deep[] holds (the column-sharded prover: partial sums over the columns a rank owns).
__global__ void __launch_bounds__(256) k_deep(const fe *__restrict__ tlde, const fe *__restrict__ clde, size_t N, int n, int blowup, int W, int C,
                                              const fe *__restrict__ dc, const fe *__restrict__ az, const fe *__restrict__ z, fe g_n,
                                              const fe *__restrict__ wpowN, const fe *__restrict__ inv, fe *__restrict__ deep, size_t s_base,
                                              size_t count, int accumulate, size_t cstride_c /* column stride of clde */,
                                              size_t c_off /* storage index of clde's first element: a coset range of the composition LDE */) {
  size_t tl = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= count) return;
  size_t s = s_base + tl;
  size_t r = (s / n) + (size_t)blowup * (s % n);
  fe x = fe_mul(fe_from_u64(3), ld_fe(wpowN + r));
  fe a = fe_zero();
  int i = 0;
  for (; i + 12 <= W; i += 12) {
    wide_acc w; wacc_zero(w);
#pragma unroll
    for (int u = 0; u < 12; u++) wacc_mad(w, ld_fe(tlde + (size_t)(i + u) * N + s), ld_fe(dc + i + u));
    a = fe_add(a, wacc_reduce(w));
  }
  for (; i < W; i++) a = fe_add(a, fe_mul(ld_fe(tlde + (size_t)i * N + s), ld_fe(dc + i)));
  for (int j = 0; j < C; j++) a = fe_add(a, fe_mul(ld_fe(clde + (size_t)j * cstride_c + (s - c_off)), ld_fe(dc + W + j)));
  fe zz = ld_fe(z);
  fe d1 = fe_sub(x, zz), d2 = fe_sub(x, fe_mul(zz, g_n));
  fe inv12 = ld_fe(inv + s);   // 1 / (d1 d2), k_deep_inverses
  fe i1 = fe_mul(inv12, d2), i2 = fe_mul(inv12, d1);
  fe y = fe_add(fe_mul(fe_sub(a, ld_fe(az)), i1), fe_mul(fe_sub(a, ld_fe(az + 1)), i2));
  if (accumulate) y = fe_add(y, ld_fe(deep + tl));
  st_fe(deep + tl, y);
}

// ---------------------------------------------------------------- FRI (agg/trace.rs:803-820, winter-fri [WF-mem])
// layer d: evaluations ev[k][i'] (coset-major, nd rows per coset, natural index k + B*i'), domain size Nd = B*nd.
// leaves: leaf[i] = hash_elements([ev[i], ev[i + Nd/2]]) for natural i < Nd/2  (transpose_slice + hash per row)
__global__ void __launch_bounds__(ZKL_COOP_THREADS) k_fri_leaves(const fe *__restrict__ ev, int nd, int B, fe *__restrict__ leaves) {
  size_t half_rows = (size_t)nd / 2, total = half_rows * B;
  size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / ZKL_COOP_GROUP;
  size_t t = g < total ? g : total - 1;
  size_t k = t / half_rows, ip = t % half_rows;
  // the two values are nd/2 apart inside coset k
  fe d = coop_hash_elements(ev + k * nd + ip, half_rows, 2);
  if (g < total && (threadIdx.x & (ZKL_COOP_GROUP - 1)) == 0) st_fe(leaves + (k + (size_t)B * ip), d);
}
// next[k][i'] = (v0 (a + x0) - v1 (a - x0)) / (2 x0), x0 = g * w_N^((k + B i') << depth)   [x1 = -x0]
// `ev` holds B cosets; they are cosets kofs .. kofs + B - 1 of a layer with Bg cosets (the whole layer: kofs = 0, Bg = B; the
// column-sharded prover folds the cosets a rank owns)
__global__ void __launch_bounds__(256) k_fri_fold(const fe *__restrict__ ev, int nd, int B, int depth, const fe *__restrict__ alpha,
                                                  const fe *__restrict__ wpowN, size_t N, fe g_inv_half /* (2g)^-1 */, fe *__restrict__ out,
                                                  int kofs, int Bg) {
  size_t half_rows = (size_t)nd / 2, total = half_rows * B;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  size_t k = t / half_rows, ip = t % half_rows;
  size_t e = ((k + kofs + (size_t)Bg * ip) << depth) & (N - 1);
  fe w = ld_fe(wpowN + e), winv = ld_fe(wpowN + (e ? N - e : 0));
  fe x0 = fe_mul(fe_from_u64(3), w);
  fe a = ld_fe(alpha);
  fe v0 = ld_fe(ev + k * nd + ip), v1 = ld_fe(ev + k * nd + ip + half_rows);
  fe num = fe_sub(fe_mul(v0, fe_add(a, x0)), fe_mul(v1, fe_sub(a, x0)));
  st_fe(out + k * half_rows + ip, fe_mul(num, fe_mul(g_inv_half, winv)));
}
// remainder (FriProver::set_remainder [WF-mem]): interpolate the last evaluations (size Nl = B*nl, offset g) and keep
// the first Nl/B = nl coefficients, written highest degree first (agg/trace.rs:941-945). Single thread, Nl <= 256.
__global__ void k_fri_remainder(const fe *__restrict__ ev, int nl, int B, const fe *__restrict__ wpowN, size_t N, fe g_inv, fe nl_inv_total,
                                fe *__restrict__ rem_rev) {
  int m = threadIdx.x;
  if (m >= nl || blockIdx.x) return;
  int Nl = nl * B;
  size_t stride = N / Nl;  // w_Nl = w_N^stride
  fe acc = fe_zero();
  for (int r = 0; r < Nl; r++) {
    int k = r % B, ip = r / B;
    size_t e = ((size_t)r * m * stride) % N;
    acc = fe_add(acc, fe_mul(ld_fe(ev + (size_t)k * nl + ip), ld_fe(wpowN + (e ? N - e : 0))));
  }
  fe c = fe_mul(fe_mul(acc, nl_inv_total), fe_pow(g_inv, (uint64_t)m));
  st_fe(rem_rev + (nl - 1 - m), c);
}
