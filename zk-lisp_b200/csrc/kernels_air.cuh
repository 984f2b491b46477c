// kernels_air.cuh — ZkLispAir transition + boundary constraint evaluation over the CE domain,
// merged into the composition polynomial trace (hook `new_evaluator` + ConstraintEvaluator::evaluate,
// zk-lisp-proof-winterfell/src/prove.rs:476-492).
//
// Constraint formulas and their ORDER follow `evaluate_transition` (vm/air/mod.rs:324-378) and the
// module `eval_block`s: poseidon.rs:67-162, ctrl.rs:114-390, alu.rs:108-354, ram.rs:82-236,
// merkle.rs:66-134, rom.rs:57-120, mixers.rs:19-48. Coefficient i multiplies constraint i.
//
// Domain: the CE domain (size ce*n, coset GENERATOR*<w_{ce n}>) is every (blowup/ce)-th LDE row. With the
// coset-major LDE layout, CE coset k' is LDE coset k = k'*(blowup/ce): one thread per (k', i) reads
// cur = lde[c][k*n + i] and next = lde[c][k*n + (i+1) mod n] — unit stride across the warp.
//
// Boundary constraints (get_assertions, vm/air/mod.rs:380-504) are NOT evaluated as Winterfell does
// (one divisor column per asserted step: O(n^2)); we use the exact identity, for a in the trace domain H,
//     1/(x - a) = n * L_a(x) / (a * (x^n - 1)),
// so  sum_k cc_k (T_{c_k}(x) - v_k)/(x - a_k) = n/(x^n-1) * ( sum_c T_c(x) P_c(x) - P_S(x) )
// with P_c, P_S the degree-<n interpolants of the sparse vectors cc_k/a_k (per asserted column) and
// cc_k v_k/a_k. Field arithmetic is exact, so the result is bit-identical (SURVEY.md §7 hard part 2).
#pragma once
#include "kernels_ntt.cuh"

struct air_dev {
  zkl_columns cols;
  uint32_t f_poseidon, f_vm, f_sponge, f_merkle, f_ram, rom_enabled;
  uint32_t vm_usage_mask, ram_delta_clk_bits;
  uint32_t num_transition;
  uint32_t air_kind;        // ZKL_AIR_ZKLISP / ZKL_AIR_AGG
  fe pose_mds[12][12];
  fe pose_rc[ZKL_ROUNDS][12];
  fe rom_mds[3][3];
  fe rom_rc[ZKL_ROUNDS][3];
  fe rom_w0[59], rom_w1[59];
  fe program_fe0, merkle_root_fe;
};

// periodic values table: per[(k' * 32 + pos) * 31 + col], the 31 cycle-32 columns of vm/air/mod.rs:520-562
// evaluated at x^(n/32) for x = GENERATOR * w_M^(k' + ce*i), which only depends on (k', i mod 32).
// Column c (c < 31) is the indicator of a set S_c of positions in the size-32 domain <w32>; its interpolant at
// y is sum_{p in S_c} L_p(y), L_p(y) = (w32^p / 32) * (y^32 - 1) / (y - w32^p).
__global__ void k_periodic_table(fe *per, int ce, int n, const fe *__restrict__ wpowM, size_t M, fe g_pow /* g^(n/32) */, fe inv32) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ce * 32) return;
  int kp = t / 32, pos = t % 32;
  // y = x^(n/32) = g^(n/32) * w_M^((kp + ce*pos) * n/32): exponent taken mod M
  size_t e = ((size_t)(kp + (size_t)ce * pos) * (size_t)(n / 32)) % M;
  fe y = fe_mul(g_pow, ld_fe(wpowM + e));
  fe y32 = y;
  for (int i = 0; i < 5; i++) y32 = fe_sqr(y32);
  fe num = fe_mul(fe_sub(y32, fe_one()), inv32);
  // w32 = w_M^(M/32)
  fe L[32];
  for (int p = 0; p < 32; p++) {
    fe wp = ld_fe(wpowM + (size_t)p * (M / 32));
    L[p] = fe_mul(fe_mul(num, wp), fe_inv(fe_sub(y, wp)));
  }
  fe *o = per + (size_t)t * 31;
  o[0] = L[0];                                    // p_map
  for (int j = 0; j < ZKL_ROUNDS; j++) o[1 + j] = L[1 + j];  // p_round[j]
  o[28] = L[28];                                  // p_final
  o[29] = fe_add(fe_add(L[29], L[30]), L[31]);    // p_pad
  o[30] = L[31];                                  // p_pad_last
}

struct frame_reader {
  const fe *cur_base;   // lde + k*n + i
  const fe *next_base;  // lde + k*n + (i+1)%n
  size_t N;
  __device__ __forceinline__ fe cur(uint32_t c) const { return ld_fe(cur_base + (size_t)c * N); }
  __device__ __forceinline__ fe next(uint32_t c) const { return ld_fe(next_base + (size_t)c * N); }
};

// sum_i cc[i] * v_i: products are accumulated UNREDUCED (wide_acc holds 12 products + one addend) and reduced once per
// 12 constraints — one reduction instead of twelve; the total is exact, so the result is the same canonical element.
struct cc_acc {
  const fe *cc; uint32_t ix; fe acc; wide_acc w; int pending;
  __device__ __forceinline__ void emit(const fe &v) {
    wacc_mad(w, ld_fe(cc + ix), v); ix++;
    if (++pending == 12) flush();
  }
  __device__ __forceinline__ void flush() {
    if (pending) { wacc_add_fe(w, acc); acc = wacc_reduce(w); wacc_zero(w); pending = 0; }
  }
};

#define MUL fe_mul
#define ADD fe_add
#define SUB fe_sub

// Sum_i cc[i] * constraint_i(frame) — transition constraints only
__device__ __noinline__ fe eval_transition_combined(const air_dev *__restrict__ air, const frame_reader &f, const fe *per31, fe p_last,
                                                    const fe *__restrict__ cc) {
  const zkl_columns &c = air->cols;
  cc_acc out; out.cc = cc; out.ix = 0; out.acc = fe_zero(); wacc_zero(out.w); out.pending = 0;
  const fe one = fe_one(), zero = fe_zero();
  fe p_map = per31[0], p_final = per31[28], p_pad = per31[29], p_pad_last = per31[30];
  const bool sponge_used = (air->vm_usage_mask >> 7) & 1;
  enum { CONST_, MOV, ADD_, SUB_, MUL_, NEG, EQ, SEL, SPONGE, ASSERT, ASSERT_BIT, ASSERT_RANGE, DIVMOD, DIV128, MULWIDE, LOAD, STORE };
#define OP(k) f.cur(c.op_start + (k))

  // ---------------- PoseidonAir (poseidon.rs:83-161). MDS*s^3 is shared by all 27 round rows.
  if (air->f_poseidon) {
    fe pa = f.cur(c.pose_active);
    fe s3[12];
#pragma unroll 1
    for (int i = 0; i < 12; i++) s3[i] = fe_cube(f.cur(c.lanes_start + i));
    fe d[12];  // next_i - (MDS*s^3)_i
#pragma unroll 1
    for (int i = 0; i < 12; i++) {
      wide_acc w; wacc_zero(w);
#pragma unroll
      for (int k = 0; k < 12; k++) wacc_mad(w, s3[k], ld_fe(&air->pose_mds[i][k]));
      d[i] = SUB(f.next(c.lanes_start + i), wacc_reduce(w));
    }
#pragma unroll 1
    for (int j = 0; j < ZKL_ROUNDS; j++) {
      fe g = MUL(pa, per31[1 + j]);
#pragma unroll 1
      for (int i = 0; i < 12; i++) out.emit(MUL(g, SUB(d[i], ld_fe(&air->pose_rc[j][i]))));
    }
    fe g_hold = SUB(p_pad, p_pad_last);
#pragma unroll 1
    for (int i = 0; i < 12; i++) out.emit(MUL(g_hold, SUB(f.next(c.lanes_start + i), f.cur(c.lanes_start + i))));
    if (air->f_vm && air->f_sponge && sponge_used) {
      fe gate = MUL(MUL(p_map, pa), OP(SPONGE));
      fe r[8];
#pragma unroll 1
      for (int i = 0; i < 8; i++) r[i] = f.cur(c.r_start + i);
#pragma unroll 1
      for (int lane = 0; lane < 10; lane++) {
        fe b0 = f.cur(c.sel_s_bits_start + lane * 3 + 0), b1 = f.cur(c.sel_s_bits_start + lane * 3 + 1),
           b2 = f.cur(c.sel_s_bits_start + lane * 3 + 2), act = f.cur(c.sel_s_active_start + lane);
        fe nb0 = SUB(one, b0), nb1 = SUB(one, b1), nb2 = SUB(one, b2);
        fe s0 = ADD(MUL(b0, r[1]), MUL(nb0, r[0])), s1 = ADD(MUL(b0, r[3]), MUL(nb0, r[2]));
        fe s2 = ADD(MUL(b0, r[5]), MUL(nb0, r[4])), s3_ = ADD(MUL(b0, r[7]), MUL(nb0, r[6]));
        fe t0 = ADD(MUL(b1, s1), MUL(nb1, s0)), t1 = ADD(MUL(b1, s3_), MUL(nb1, s2));
        fe sel_val = ADD(MUL(b2, t1), MUL(nb2, t0));
        out.emit(MUL(gate, SUB(f.cur(c.lanes_start + lane), MUL(act, sel_val))));
      }
    }
  }

  fe s_low = MUL(p_last, p_map);  // mixers.rs:19-25
  fe g_carry = ADD(p_map, SUB(p_pad, p_pad_last));
#pragma unroll 1
  for (int j = 0; j < ZKL_ROUNDS - 1; j++) g_carry = ADD(g_carry, per31[1 + j]);

  if (air->f_vm) {
    // ---------------- VmCtrlAir (ctrl.rs:162-389)
    fe pi_prog = f.cur(c.pi_prog);
    fe s_high = MUL(s_low, pi_prog);
    fe sum_dst0 = zero, sum_a = zero, sum_b = zero, sum_c = zero, sum_dst1 = zero;
#pragma unroll 1
    for (int i = 0; i < ZKL_NR; i++) {
      fe sd0 = f.cur(c.sel_dst0_start + i), sa = f.cur(c.sel_a_start + i), sb = f.cur(c.sel_b_start + i),
         sc = f.cur(c.sel_c_start + i), sd1 = f.cur(c.sel_dst1_start + i);
      sum_dst0 = ADD(sum_dst0, sd0); sum_a = ADD(sum_a, sa); sum_b = ADD(sum_b, sb); sum_c = ADD(sum_c, sc); sum_dst1 = ADD(sum_dst1, sd1);
      out.emit(ADD(MUL(MUL(p_map, sd0), SUB(sd0, one)), s_high));
      out.emit(ADD(MUL(MUL(p_map, sa), SUB(sa, one)), s_high));
      out.emit(ADD(MUL(MUL(p_map, sb), SUB(sb, one)), s_high));
      out.emit(ADD(MUL(MUL(p_map, sc), SUB(sc, one)), s_high));
      out.emit(ADD(MUL(MUL(p_map, sd1), SUB(sd1, one)), s_high));
    }
    fe ops[ZKL_NUM_OPS];
#pragma unroll 1
    for (int k = 0; k < ZKL_NUM_OPS; k++) ops[k] = OP(k);
    fe uses_a = ADD(ADD(ADD(ADD(ADD(ADD(ops[MOV], ops[ADD_]), ops[SUB_]), ops[MUL_]), ops[NEG]), ops[EQ]), ops[SEL]);
    uses_a = ADD(ADD(ADD(ADD(ADD(uses_a, ops[DIVMOD]), ops[DIV128]), ops[MULWIDE]), ops[LOAD]), ops[STORE]);
    fe uses_b = ADD(ADD(ADD(ADD(ops[ADD_], ops[SUB_]), ops[MUL_]), ops[EQ]), ops[SEL]);
    uses_b = ADD(ADD(ADD(ADD(uses_b, ops[DIVMOD]), ops[DIV128]), ops[MULWIDE]), ops[STORE]);
    fe uses_c = ADD(ADD(ADD(ops[SEL], ops[ASSERT]), ops[ASSERT_BIT]), ops[ASSERT_RANGE]);
    fe op_any = zero;
#pragma unroll 1
    for (int k = 0; k <= MULWIDE; k++) op_any = ADD(op_any, ops[k]);
    fe uses_dst0 = ADD(SUB(op_any, ops[SPONGE]), ops[LOAD]);
    fe uses_dst1 = ADD(ADD(ops[DIVMOD], ops[DIV128]), ops[MULWIDE]);
    out.emit(ADD(MUL(p_map, SUB(sum_dst0, uses_dst0)), s_low));
    out.emit(ADD(MUL(p_map, SUB(sum_a, uses_a)), s_low));
    out.emit(ADD(MUL(p_map, SUB(sum_b, uses_b)), s_low));
    out.emit(ADD(MUL(p_map, SUB(sum_c, uses_c)), s_low));
    out.emit(ADD(MUL(p_map, SUB(sum_dst1, uses_dst1)), s_low));
#pragma unroll 1
    for (int i = 0; i < ZKL_NR; i++) out.emit(ADD(MUL(MUL(p_map, f.cur(c.sel_dst0_start + i)), f.cur(c.sel_dst1_start + i)), s_high));
    if (air->f_sponge && sponge_used) {
#pragma unroll 1
      for (int lane = 0; lane < 10; lane++) {
#pragma unroll 1
        for (int b = 0; b < 3; b++) {
          fe bv = f.cur(c.sel_s_bits_start + lane * 3 + b);
          out.emit(ADD(MUL(MUL(p_map, bv), SUB(bv, one)), s_high));
        }
        fe act = f.cur(c.sel_s_active_start + lane);
        out.emit(ADD(MUL(MUL(p_map, act), SUB(act, one)), s_high));
      }
    }
    out.emit(s_high);  // ctrl.rs:292-293
    fe op_sum = ADD(ADD(op_any, ops[LOAD]), ops[STORE]);
#pragma unroll 1
    for (int k = 0; k < ZKL_NUM_OPS; k++) out.emit(ADD(MUL(MUL(p_map, ops[k]), SUB(ops[k], one)), s_high));
    out.emit(ADD(MUL(MUL(p_map, op_sum), SUB(op_sum, one)), s_high));
#pragma unroll 1
    for (int k = 0; k < ZKL_NUM_OPS; k++) {
      fe v = air->rom_enabled ? MUL(p_map, SUB(ops[k], f.cur(c.rom_op_start + k))) : zero;
      out.emit(ADD(v, s_high));
    }
    fe pc_cur = f.cur(c.pc), pc_next = f.next(c.pc);
    out.emit(ADD(air->rom_enabled ? MUL(g_carry, SUB(pc_next, pc_cur)) : zero, s_low));
    out.emit(ADD(air->rom_enabled ? MUL(p_pad_last, SUB(pc_next, ADD(pc_cur, one))) : zero, s_low));

    // ---------------- VmAluAir (alu.rs:162-353)
    const uint32_t m = air->vm_usage_mask;
    const bool use_assert = m & 1, use_assert_bit = (m >> 1) & 1, use_assert_range = (m >> 2) & 1, use_divmod = (m >> 3) & 1,
               use_mulwide = (m >> 4) & 1, use_div128 = (m >> 5) & 1, use_eq = (m >> 6) & 1;
    fe pi2 = fe_sqr(pi_prog), pi4 = fe_sqr(pi2), pi6 = MUL(pi4, pi2);
    fe s_write = MUL(s_low, pi6), s_eq = MUL(s_low, pi4);
    fe a_val = zero, b_val = zero, c_val = zero, dst0_next = zero, dst0_cur = zero, dst1_next = zero;
#pragma unroll 1
    for (int i = 0; i < ZKL_NR; i++) {
      fe r = f.cur(c.r_start + i), rn = f.next(c.r_start + i);
      fe sd0 = f.cur(c.sel_dst0_start + i);
      a_val = ADD(a_val, MUL(f.cur(c.sel_a_start + i), r));
      b_val = ADD(b_val, MUL(f.cur(c.sel_b_start + i), r));
      c_val = ADD(c_val, MUL(f.cur(c.sel_c_start + i), r));
      dst0_next = ADD(dst0_next, MUL(sd0, rn));
      dst0_cur = ADD(dst0_cur, MUL(sd0, r));
      dst1_next = ADD(dst1_next, MUL(f.cur(c.sel_dst1_start + i), rn));
      out.emit(ADD(MUL(g_carry, SUB(rn, r)), s_low));  // carry constraints are emitted first (alu.rs:176-179)
    }
    fe imm = f.cur(c.imm), eq_inv = f.cur(c.eq_inv);
    fe res = MUL(ops[CONST_], imm);
    res = ADD(res, MUL(ops[MOV], a_val));
    res = ADD(res, MUL(ops[ADD_], ADD(a_val, b_val)));
    res = ADD(res, MUL(ops[SUB_], SUB(a_val, b_val)));
    res = ADD(res, MUL(ops[MUL_], MUL(a_val, b_val)));
    res = ADD(res, MUL(ops[NEG], SUB(zero, a_val)));
    res = ADD(res, MUL(ops[SEL], ADD(MUL(c_val, a_val), MUL(SUB(one, c_val), b_val))));
    res = ADD(res, MUL(ops[SPONGE], f.cur(c.lanes_start)));
    if (use_eq) res = ADD(res, MUL(ops[EQ], dst0_next));
    if (use_assert) res = ADD(res, ops[ASSERT]);
    if (use_assert_bit) res = ADD(res, ops[ASSERT_BIT]);
    res = ADD(res, MUL(ops[LOAD], imm));
    fe sum = zero, pow2 = one;
#pragma unroll 1
    for (int i = 0; i < 32; i++) { sum = ADD(sum, MUL(pow2, f.cur(c.gadget_b_start + i))); pow2 = ADD(pow2, pow2); }
    if (use_assert_range) res = ADD(res, MUL(ops[ASSERT_RANGE], ADD(MUL(SUB(one, imm), sum), imm)));
    const bool uses_two = use_divmod || use_mulwide || use_div128;
    fe b_two = uses_two ? ADD(ADD(ops[DIVMOD], ops[MULWIDE]), ops[DIV128]) : zero;
    fe w0 = ADD(MUL(SUB(one, b_two), res), MUL(b_two, dst0_next));
    fe w1 = MUL(b_two, dst1_next);
#pragma unroll 1
    for (int i = 0; i < ZKL_NR; i++) {
      fe sd0 = f.cur(c.sel_dst0_start + i), sd1 = f.cur(c.sel_dst1_start + i);
      fe keep = SUB(SUB(one, sd0), sd1);
      fe rhs = ADD(ADD(MUL(keep, f.cur(c.r_start + i)), MUL(sd0, w0)), MUL(sd1, w1));
      out.emit(ADD(MUL(p_final, SUB(f.next(c.r_start + i), rhs)), s_write));
    }
    fe diff = SUB(a_val, b_val);
    if (use_eq) {
      fe g = MUL(p_final, ops[EQ]);
      out.emit(ADD(MUL(g, MUL(dst0_next, diff)), s_eq));
      out.emit(ADD(MUL(g, SUB(SUB(one, dst0_next), MUL(diff, eq_inv))), s_eq));
    }
    if (use_divmod) {
      fe g = MUL(p_final, ops[DIVMOD]);
      out.emit(ADD(MUL(g, SUB(SUB(a_val, MUL(b_val, dst0_next)), dst1_next)), s_eq));
      out.emit(ADD(MUL(g, SUB(MUL(b_val, eq_inv), one)), s_eq));
    }
    fe p2_64; p2_64.l[0] = 0; p2_64.l[1] = 0; p2_64.l[2] = 1; p2_64.l[3] = 0;
    if (use_mulwide) out.emit(ADD(MUL(MUL(p_final, ops[MULWIDE]), SUB(MUL(a_val, b_val), ADD(dst0_next, MUL(dst1_next, p2_64)))), s_eq));
    if (use_div128) {
      fe g = MUL(p_final, ops[DIV128]);
      fe num128 = ADD(MUL(a_val, p2_64), imm);
      out.emit(ADD(MUL(g, SUB(num128, ADD(MUL(b_val, dst0_next), dst1_next))), s_eq));
      out.emit(ADD(MUL(g, SUB(MUL(b_val, eq_inv), one)), s_eq));
    }
    if (use_assert) out.emit(ADD(MUL(p_final, ADD(MUL(ops[ASSERT], SUB(c_val, one)), MUL(ops[SEL], MUL(c_val, SUB(c_val, one))))), s_eq));
    if (use_assert_bit) out.emit(ADD(MUL(MUL(p_final, ops[ASSERT_BIT]), MUL(c_val, SUB(c_val, one))), s_eq));
    if (use_assert_range) {
      fe g = MUL(p_final, ops[ASSERT_RANGE]);
#pragma unroll 1
      for (int i = 0; i < 32; i++) {
        fe bi = f.cur(c.gadget_b_start + i);
        out.emit(ADD(MUL(g, MUL(bi, SUB(bi, one))), s_eq));
      }
      fe p2_32; p2_32.l[0] = 0; p2_32.l[1] = 1; p2_32.l[2] = 0; p2_32.l[3] = 0;
      fe eq32 = SUB(c_val, sum), eq64 = SUB(c_val, ADD(dst0_cur, MUL(sum, p2_32)));
      fe eq_term = MUL(imm, ADD(MUL(eq_inv, eq64), MUL(SUB(one, eq_inv), eq32)));  // mode64 = eq_inv column (alu.rs:180)
      out.emit(ADD(MUL(g, eq_term), s_eq));
    }
  }

  // ---------------- RamAir (ram.rs:137-235)
  if (air->f_ram) {
    const uint32_t R = c.ram_start;
    fe g_hold = SUB(p_pad, p_pad_last);
    fe op_load = OP(LOAD), op_store = OP(STORE);
    fe event = MUL(p_final, ADD(op_load, op_store));
    fe pi0 = ld_fe(&air->program_fe0);
    fe pi2 = fe_sqr(pi0), pi3 = MUL(pi2, pi0), pi4 = fe_sqr(pi2), pi5 = MUL(pi4, pi0);
    fe r1 = ADD(pi2, one), r2 = ADD(pi3, pi0), r3 = ADD(pi5, fe_from_u64(7));
    fe a_ev = zero, b_ev = zero;
#pragma unroll 1
    for (int i = 0; i < ZKL_NR; i++) {
      fe ri = f.cur(c.r_start + i);
      a_ev = ADD(a_ev, MUL(f.cur(c.sel_a_start + i), ri));
      b_ev = ADD(b_ev, MUL(f.cur(c.sel_b_start + i), ri));
    }
    fe w_ev = op_store;
    fe val_ev = ADD(MUL(w_ev, b_ev), MUL(SUB(one, w_ev), f.cur(c.imm)));
    fe comp_uns = ADD(ADD(ADD(a_ev, MUL(r1, f.cur(c.pc))), MUL(r2, val_ev)), MUL(r3, w_ev));
    fe gu = f.cur(R + 6), gu_n = f.next(R + 6), gs = f.cur(R + 7), gs_n = f.next(R + 7);
    fe dgu = SUB(gu_n, gu);
    out.emit(ADD(ADD(MUL(event, SUB(gu_n, ADD(gu, comp_uns))), MUL(SUB(one, event), dgu)), MUL(g_hold, dgu)));
    fe s_on = f.cur(R + 0), s_addr = f.cur(R + 1), s_clk = f.cur(R + 2), s_val = f.cur(R + 3), s_w = f.cur(R + 4), last = f.cur(R + 5);
    fe s_addr_n = f.next(R + 1), s_clk_n = f.next(R + 2), last_n = f.next(R + 5);
    fe same = SUB(one, MUL(SUB(s_addr_n, s_addr), f.cur(c.eq_inv)));
    fe comp = ADD(ADD(ADD(s_addr, MUL(r1, s_clk)), MUL(r2, s_val)), MUL(r3, s_w));
    out.emit(ADD(MUL(s_on, SUB(gs_n, ADD(gs, comp))), MUL(SUB(one, s_on), SUB(gs_n, gs))));
    fe wv = MUL(s_w, s_val);
    fe last_keep = ADD(MUL(same, ADD(MUL(SUB(one, s_w), last), wv)), MUL(SUB(one, same), wv));
    out.emit(MUL(s_on, SUB(last_n, last_keep)));
    out.emit(MUL(MUL(s_on, SUB(one, s_w)), SUB(s_val, last)));
    fe s_on_n = f.next(R + 0), s_w_n = f.next(R + 4), s_val_n = f.next(R + 3);
    out.emit(MUL(MUL(MUL(MUL(s_on, s_on_n), SUB(one, same)), SUB(one, s_w_n)), s_val_n));
    out.emit(MUL(s_on, MUL(same, SUB(same, one))));
    if ((air->vm_usage_mask >> 8) & 1) {
      fe d_clk = SUB(s_clk_n, s_clk);
      fe sum = zero, pow2 = one, g = MUL(s_on, same);
#pragma unroll 1
      for (int i = 0; i < 32; i++) {
        fe bi = f.cur(c.gadget_b_start + i);
        if ((air->ram_delta_clk_bits >> i) & 1) out.emit(MUL(MUL(g, bi), SUB(bi, one)));
        sum = ADD(sum, MUL(pow2, bi)); pow2 = ADD(pow2, pow2);
      }
      out.emit(MUL(MUL(MUL(s_on, s_on_n), same), SUB(d_clk, sum)));
    }
    out.emit(MUL(p_last, SUB(gu, gs)));
  }

  // ---------------- MerkleAir (merkle.rs:90-133)
  if (air->f_merkle) {
    const uint32_t Mk = c.merkle_start;
    fe g = f.cur(Mk + 0), dir = f.cur(Mk + 1), sib = f.cur(Mk + 2), acc = f.cur(Mk + 3);
    fe pg = MUL(p_map, g);
    out.emit(MUL(MUL(pg, dir), SUB(dir, one)));
    fe nd = SUB(one, dir);
    out.emit(MUL(pg, SUB(f.cur(c.lanes_start + 0), ADD(MUL(nd, acc), MUL(dir, sib)))));
    out.emit(MUL(pg, SUB(f.cur(c.lanes_start + 1), ADD(MUL(nd, sib), MUL(dir, acc)))));
    fe dacc = SUB(f.next(Mk + 3), acc);
    out.emit(MUL(MUL(g, g_carry), dacc));
    out.emit(MUL(MUL(pg, f.cur(Mk + 4)), SUB(acc, f.cur(Mk + 6))));
    out.emit(MUL(MUL(MUL(p_final, g), f.cur(Mk + 5)), SUB(acc, ld_fe(&air->merkle_root_fe))));
    out.emit(MUL(MUL(MUL(p_pad_last, g), f.next(Mk + 0)), dacc));
  }

  // ---------------- RomAir (rom.rs:71-119)
  if (air->rom_enabled) {
    fe s3[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; i++) s3[i] = fe_cube(f.cur(c.rom_s_start + i));
#pragma unroll
    for (int i = 0; i < 3; i++) {
      wide_acc w; wacc_zero(w);
#pragma unroll
      for (int k = 0; k < 3; k++) wacc_mad(w, s3[k], ld_fe(&air->rom_mds[i][k]));
      d[i] = SUB(f.next(c.rom_s_start + i), wacc_reduce(w));
    }
#pragma unroll 1
    for (int j = 0; j < ZKL_ROUNDS; j++) {
      fe gr = per31[1 + j];
#pragma unroll
      for (int i = 0; i < 3; i++) out.emit(MUL(gr, SUB(d[i], ld_fe(&air->rom_rc[j][i]))));
    }
    fe g_hold = SUB(p_pad, p_pad_last);
#pragma unroll
    for (int i = 0; i < 3; i++) out.emit(MUL(g_hold, SUB(f.next(c.rom_s_start + i), f.cur(c.rom_s_start + i))));
    fe enc0 = zero, enc1 = zero;
    int k = 0;
#pragma unroll 1
    for (int o = 0; o < ZKL_NUM_OPS; o++, k++) {
      fe v = f.cur(c.op_start + o);
      enc0 = ADD(enc0, MUL(v, ld_fe(&air->rom_w0[k]))); enc1 = ADD(enc1, MUL(v, ld_fe(&air->rom_w1[k])));
    }
    const uint32_t starts[5] = {c.sel_dst0_start, c.sel_a_start, c.sel_b_start, c.sel_c_start, c.sel_dst1_start};
#pragma unroll 1
    for (int g = 0; g < 5; g++) {
#pragma unroll 1
      for (int i = 0; i < ZKL_NR; i++, k++) {
        fe v = f.cur(starts[g] + i);
        enc0 = ADD(enc0, MUL(v, ld_fe(&air->rom_w0[k]))); enc1 = ADD(enc1, MUL(v, ld_fe(&air->rom_w1[k])));
      }
    }
    out.emit(MUL(p_map, SUB(f.cur(c.rom_s_start + 1), enc0)));
    out.emit(MUL(p_map, SUB(f.cur(c.rom_s_start + 2), enc1)));
  }
#undef OP
  out.flush();
  return out.acc;
}
#undef MUL
#undef ADD
#undef SUB

// ZlAggAir::evaluate_transition (agg/air.rs:112-272; columns agg/layout.rs:97-180), combined with the coefficients: 24 constraints
// over 31 columns, `p_last` is the AIR's one periodic column (is_last, cycle n).
__device__ __noinline__ fe eval_transition_combined_agg(const frame_reader &f, fe p_last, const fe *__restrict__ cc) {
  enum { OK = 0, V0_SUM, V1_SUM, VNEXT_SUM, FRI_V0, FRI_V1, FRI_VNEXT, FRI_ALPHA, FRI_X0, FRI_X1, FRI_Q1, COMP_SUM, ALPHA_DIV_ZM, MAP_L0, FINAL_LLAST,
         R_, ALPHA, BETA, GAMMA, SEG_FIRST, TRACE_ROOT_ERR, CONS_ROOT_ERR, V_UNITS_ACC, V_UNITS_CHILD, CHILD_COUNT_ACC, VM_ERR, RAM_U_ERR, RAM_S_ERR,
         ROM_ERR0, ROM_ERR1, ROM_ERR2 };
  cc_acc out; out.cc = cc; out.ix = 0; out.acc = fe_zero(); wacc_zero(out.w); out.pending = 0;
  const fe not_last = fe_sub(fe_one(), p_last);
  out.emit(f.cur(OK));
  out.emit(fe_mul(not_last, fe_sub(f.next(V_UNITS_ACC), fe_add(f.cur(V_UNITS_ACC), fe_mul(f.cur(V_UNITS_CHILD), f.cur(SEG_FIRST))))));
  out.emit(f.cur(TRACE_ROOT_ERR));
  out.emit(f.cur(CONS_ROOT_ERR));
  const int chained[7] = {R_, ALPHA, BETA, GAMMA, V0_SUM, V1_SUM, VNEXT_SUM};
#pragma unroll
  for (int k = 0; k < 7; k++) out.emit(fe_mul(not_last, fe_sub(f.next(chained[k]), f.cur(chained[k]))));
  out.emit(fe_mul(not_last, fe_sub(f.next(CHILD_COUNT_ACC), fe_add(f.cur(CHILD_COUNT_ACC), f.cur(SEG_FIRST)))));
  const fe x0 = f.cur(FRI_X0), x1 = f.cur(FRI_X1), al = f.cur(FRI_ALPHA), vn = f.cur(FRI_VNEXT);
  const fe rhs = fe_sub(fe_mul(f.cur(FRI_V1), fe_sub(al, x0)), fe_mul(f.cur(FRI_V0), fe_sub(al, x1)));
  out.emit(fe_sub(fe_mul(vn, fe_sub(x1, x0)), rhs));
  out.emit(fe_sub(vn, f.cur(FRI_Q1)));
  const int zero_cols[10] = {COMP_SUM, ALPHA_DIV_ZM, MAP_L0, FINAL_LLAST, VM_ERR, RAM_U_ERR, RAM_S_ERR, ROM_ERR0, ROM_ERR1, ROM_ERR2};
#pragma unroll
  for (int k = 0; k < 10; k++) out.emit(f.cur(zero_cols[k]));
  out.flush();
  return out.acc;
}

// boundary data for the kernel: `slots` asserted columns, bpoly[(slot) * M + s] = P_c over the CE domain
// (coset-major, M = ce*n), slot == nslots is P_S.
struct boundary_dev {
  const uint32_t *slot_col;  // nslots entries: trace column of each slot
  const fe *bpoly;           // (nslots + 1) x M, or (nslots + 1) x (ncos * n) for the cosets of the call when `compact`
  int nslots;
  size_t stride; int compact;
};

// p_last table (data independent, cached per (n, ce)): plast[k'*n + i] = L_{n-1}(x) = g_last (x^n - 1) / (n (x - g_last)) at the
// CE point x = g w_N^(k + blowup i) — takes the per-point field inversion out of k_eval_constraints.
__global__ void k_plast_table(const fe *__restrict__ wpowN, size_t N, int n, int blowup, int ce, const fe *__restrict__ coset_consts, fe *__restrict__ plast) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)t >= (size_t)ce * n) return;
  int kp = t / n, i = t - kp * n, k = kp * (blowup / ce);
  fe x = fe_mul(fe_from_u64(3), ld_fe(wpowN + ((size_t)k + (size_t)blowup * i)));
  fe g_last = ld_fe(wpowN + (N - (size_t)blowup));
  fe xn1 = ld_fe(coset_consts + 2 * kp + 1), ninv = ld_fe(coset_consts + 2 * ce);
  st_fe(plast + t, fe_mul(fe_mul(fe_mul(g_last, xn1), ninv), fe_inv(fe_sub(x, g_last))));
}

// ce_out[k'*n + i] = transition/Z + boundary   (coset-major over the CE domain)
// The call covers CE cosets kp0 .. kp0 + ncos - 1. `lde` holds them `cos_step * n` elements apart with column stride `cstride`:
// the whole coset-major LDE (kp0 = 0, ncos = ce, cos_step = blowup / ce, cstride = N) or, in the column-sharded prover, the rows a
// rank received for ITS cosets (cos_step = 1, cstride = ncos * n). Tables and the output are indexed by the global (kp, i).
__global__ void __launch_bounds__(128) k_eval_constraints(const air_dev *__restrict__ air, const fe *__restrict__ lde, size_t N, int n, int blowup,
                                                          int ce, const fe *__restrict__ per, const fe *__restrict__ wpowN,
                                                          const fe *__restrict__ cc, boundary_dev bd, const fe *__restrict__ coset_consts,
                                                          const fe *__restrict__ plast, fe *__restrict__ ce_out, size_t cstride, int kp0, int ncos,
                                                          int cos_step) {
  int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)tl >= (size_t)ncos * n) return;
  int j = tl / n, i = tl - j * n, kp = kp0 + j;
  int t = kp * n + i;
  int step = blowup / ce, k = kp * step;
  const fe *cbase = lde + (size_t)j * cos_step * n;
  frame_reader f{cbase + i, cbase + ((i + 1) & (n - 1)), cstride};
  // x = g * w_N^(k + blowup*i)
  fe x = fe_mul(fe_from_u64(3), ld_fe(wpowN + ((size_t)k + (size_t)blowup * i)));
  // per-coset constants: [0] = (x^n - 1)^-1, [1] = x^n - 1 ; global: g_n^-1 = w_N^(N - blowup), n^-1
  fe zinv = ld_fe(coset_consts + 2 * kp), xn1 = ld_fe(coset_consts + 2 * kp + 1);
  fe g_last = ld_fe(wpowN + (N - (size_t)blowup));           // g_n^(n-1)
  fe nfe = ld_fe(coset_consts + 2 * ce + 1);
  fe xm = fe_sub(x, g_last);
  // p_last(x) = L_{n-1}(x) = g_last * (x^n - 1) / (n * (x - g_last))   (k_plast_table)
  fe p_last = ld_fe(plast + t);
  (void)xn1;
  fe tr;
  if (air->air_kind == ZKL_AIR_AGG) tr = eval_transition_combined_agg(f, p_last, cc);
  else {
    const fe *per31 = per + ((size_t)kp * 32 + (i & 31)) * 31;
    tr = eval_transition_combined(air, f, per31, p_last, cc);
  }
  // transition divisor z(x) = (x^n - 1)/(x - g^(n-1))
  fe res = fe_mul(fe_mul(tr, xm), zinv);
  // boundary part: n/(x^n-1) * (sum_c T_c(x) P_c(x) - P_S(x))
  fe b = fe_zero();
  const size_t bt = bd.compact ? (size_t)tl : (size_t)t;
  for (int s0 = 0; s0 < bd.nslots; s0 += 12) {   // unreduced accumulation, 12 products per reduction
    wide_acc w; wacc_zero(w);
    int s1 = min(bd.nslots, s0 + 12);
    for (int s = s0; s < s1; s++) wacc_mad(w, f.cur(bd.slot_col[s]), ld_fe(bd.bpoly + (size_t)s * bd.stride + bt));
    wacc_add_fe(w, b);
    b = wacc_reduce(w);
  }
  b = fe_sub(b, ld_fe(bd.bpoly + (size_t)bd.nslots * bd.stride + bt));
  res = fe_add(res, fe_mul(fe_mul(b, nfe), zinv));
  st_fe(ce_out + t, res);
}

// scatter the sparse boundary vectors: one thread per asserted step (group). For assertion q (sorted order)
// at (slot, step, value): u[slot][step] = cc_q * w_n^-step ; uS[step] += cc_q * value * w_n^-step.
struct assertion_dev { uint32_t slot, step, cc_index, pad; fe value; };
__global__ void k_boundary_scatter(const assertion_dev *__restrict__ as, const uint32_t *__restrict__ group_start, int ngroups,
                                   const fe *__restrict__ cc, const fe *__restrict__ wpow_n, int n, int nslots, fe *__restrict__ u) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  uint32_t b = group_start[g], e = group_start[g + 1];
  uint32_t step = as[b].step;
  fe ainv = ld_fe(wpow_n + ((n - step) & (n - 1)));
  fe s = fe_zero();
  for (uint32_t q = b; q < e; q++) {
    fe w = fe_mul(ld_fe(cc + as[q].cc_index), ainv);
    st_fe(u + (size_t)as[q].slot * n + step, w);
    s = fe_add(s, fe_mul(w, ld_fe(&as[q].value)));
  }
  st_fe(u + (size_t)nslots * n + step, s);
}
