// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/f128.h header).
// CPU restatement of ZkLispAir::evaluate_transition and the periodic columns:
//   /root/reference/zk-lisp-proof-winterfell/src/vm/air/mod.rs:324-378,520-592
//   .../vm/air/poseidon.rs:67-162   .../vm/air/ctrl.rs:114-390   .../vm/air/alu.rs:108-354
//   .../vm/air/ram.rs:82-236        .../vm/air/merkle.rs:66-134  .../vm/air/rom.rs:57-120
//   .../vm/air/mixers.rs:19-48      .../vm/schedule.rs:18-32
// The constraint index order is the one `eval_block` writes (coefficient i multiplies result[i]).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/zkl_b200.h"
#include "stark_core.h"

namespace zkl {

static inline F FE(const zkl_fe &x) { F r; r.v = mk128(x.hi, x.lo); return r; }

// periodic column indices (vm/air/mod.rs:520-562)
enum { P_MAP = 0, P_ROUND0 = 1, P_FINAL = 1 + ROUNDS, P_PAD = 2 + ROUNDS, P_PAD_LAST = 3 + ROUNDS, P_LAST = 4 + ROUNDS };

struct PeriodicPolys {
  size_t n;
  std::vector<F> cyc[31];  // 31 polys over the size-32 domain
  std::vector<F> last;     // p_last over the size-n domain
  explicit PeriodicPolys(size_t n_) : n(n_) {
    for (int k = 0; k < 31; k++) cyc[k].assign(32, f_zero());
    for (int pos = 0; pos < 32; pos++) {
      if (pos == 0) cyc[P_MAP][pos] = f_one();
      for (int j = 0; j < ROUNDS; j++) if (pos == 1 + j) cyc[P_ROUND0 + j][pos] = f_one();
      if (pos == 1 + ROUNDS) cyc[P_FINAL][pos] = f_one();
      bool is_pad = pos != 0 && pos != 1 + ROUNDS && !(pos >= 1 && pos <= ROUNDS);
      if (is_pad) cyc[P_PAD][pos] = f_one();
      if (pos == 31) cyc[P_PAD_LAST][pos] = f_one();
    }
    for (int k = 0; k < 31; k++) interpolate(cyc[k]);
    last.assign(n, f_zero()); last[n - 1] = f_one(); interpolate(last);
  }
  // Winterfell: a periodic poly of length L is evaluated at x^(n/L)  [WF-mem; src/preflight.rs:116-120]
  void eval_at(F x, F out[32]) const {
    F xc = f_pow(x, n / 32);
    for (int k = 0; k < 31; k++) out[k] = poly_eval(cyc[k].data(), 32, xc);
    out[P_LAST] = poly_eval(last.data(), n, x);
  }
};

static inline F sq(F a) { return f_mul(a, a); }

// ZlAggAir::evaluate_transition (agg/air.rs:112-272, columns agg/layout.rs:97-180): 24 values; per[P_LAST] is the AIR's one periodic
// column (is_last, cycle n).
static inline void eval_transition_agg(const F *cur, const F *next, const F *per, F *r) {
  enum { OK = 0, V0_SUM, V1_SUM, VNEXT_SUM, FRI_V0, FRI_V1, FRI_VNEXT, FRI_ALPHA, FRI_X0, FRI_X1, FRI_Q1, COMP_SUM, ALPHA_DIV_ZM, MAP_L0, FINAL_LLAST,
         R_, ALPHA, BETA, GAMMA, SEG_FIRST, TRACE_ROOT_ERR, CONS_ROOT_ERR, V_UNITS_ACC, V_UNITS_CHILD, CHILD_COUNT_ACC, VM_ERR, RAM_U_ERR, RAM_S_ERR,
         ROM_ERR0, ROM_ERR1, ROM_ERR2 };
  const F not_last = f_sub(f_one(), per[P_LAST]);
  r[0] = cur[OK];
  r[1] = f_mul(not_last, f_sub(next[V_UNITS_ACC], f_add(cur[V_UNITS_ACC], f_mul(cur[V_UNITS_CHILD], cur[SEG_FIRST]))));
  r[2] = cur[TRACE_ROOT_ERR];
  r[3] = cur[CONS_ROOT_ERR];
  const int chained[7] = {R_, ALPHA, BETA, GAMMA, V0_SUM, V1_SUM, VNEXT_SUM};
  for (int k = 0; k < 7; k++) r[4 + k] = f_mul(not_last, f_sub(next[chained[k]], cur[chained[k]]));
  r[11] = f_mul(not_last, f_sub(next[CHILD_COUNT_ACC], f_add(cur[CHILD_COUNT_ACC], cur[SEG_FIRST])));
  const F lhs = f_mul(cur[FRI_VNEXT], f_sub(cur[FRI_X1], cur[FRI_X0]));
  const F rhs = f_sub(f_mul(cur[FRI_V1], f_sub(cur[FRI_ALPHA], cur[FRI_X0])), f_mul(cur[FRI_V0], f_sub(cur[FRI_ALPHA], cur[FRI_X1])));
  r[12] = f_sub(lhs, rhs);
  r[13] = f_sub(cur[FRI_VNEXT], cur[FRI_Q1]);
  r[14] = cur[COMP_SUM]; r[15] = cur[ALPHA_DIV_ZM]; r[16] = cur[MAP_L0]; r[17] = cur[FINAL_LLAST];
  r[18] = cur[VM_ERR]; r[19] = cur[RAM_U_ERR]; r[20] = cur[RAM_S_ERR]; r[21] = cur[ROM_ERR0]; r[22] = cur[ROM_ERR1]; r[23] = cur[ROM_ERR2];
}

// writes air.num_transition values
static inline void eval_transition(const zkl_air_params &air, const F *cur, const F *next, const F *per, F *result) {
  if (air.air_kind == ZKL_AIR_AGG) { eval_transition_agg(cur, next, per, result); return; }
  const zkl_columns &c = air.cols;
  size_t ix = 0;
  const F one = f_one(), zero = f_zero();
  auto mul = [](F a, F b) { return f_mul(a, b); };
  auto add = [](F a, F b) { return f_add(a, b); };
  auto sub = [](F a, F b) { return f_sub(a, b); };
  F p_map = per[P_MAP], p_final = per[P_FINAL], p_pad = per[P_PAD], p_pad_last = per[P_PAD_LAST], p_last = per[P_LAST];
  bool sponge_used = (air.vm_usage_mask >> 7) & 1;
  auto OP = [&](int k) { return cur[c.op_start + k]; };
  enum { CONST_, MOV, ADD, SUB, MUL, NEG, EQ, SEL, SPONGE, ASSERT, ASSERT_BIT, ASSERT_RANGE, DIVMOD, DIV128, MULWIDE, LOAD, STORE };

  // ---------------- PoseidonAir (poseidon.rs:67-162)
  if (air.f_poseidon) {
    F pa = cur[c.pose_active];
    F s3[12];
    for (int i = 0; i < 12; i++) s3[i] = f_cube(cur[c.lanes_start + i]);
    for (int j = 0; j < ROUNDS; j++) {
      F gr = per[P_ROUND0 + j];
      for (int i = 0; i < 12; i++) {
        F acc = zero;
        for (int k = 0; k < 12; k++) acc = add(acc, mul(FE(air.pose.mds[i][k]), s3[k]));
        F y = add(acc, FE(air.pose.rc[j][i]));
        result[ix++] = mul(mul(pa, gr), sub(next[c.lanes_start + i], y));
      }
    }
    F g_hold = sub(p_pad, p_pad_last);
    for (int i = 0; i < 12; i++) result[ix++] = mul(g_hold, sub(next[c.lanes_start + i], cur[c.lanes_start + i]));
    if (air.f_vm && air.f_sponge && sponge_used) {
      F b_sponge = OP(SPONGE);
      for (int lane = 0; lane < 10; lane++) {
        F b0 = cur[c.sel_s_bits_start + lane * 3 + 0], b1 = cur[c.sel_s_bits_start + lane * 3 + 1],
          b2 = cur[c.sel_s_bits_start + lane * 3 + 2], act = cur[c.sel_s_active_start + lane];
        const F *r = cur + c.r_start;
        F s0 = add(mul(b0, r[1]), mul(sub(one, b0), r[0]));
        F s1 = add(mul(b0, r[3]), mul(sub(one, b0), r[2]));
        F s2 = add(mul(b0, r[5]), mul(sub(one, b0), r[4]));
        F s3_ = add(mul(b0, r[7]), mul(sub(one, b0), r[6]));
        F t0 = add(mul(b1, s1), mul(sub(one, b1), s0));
        F t1 = add(mul(b1, s3_), mul(sub(one, b1), s2));
        F sel_val = add(mul(b2, t1), mul(sub(one, b2), t0));
        F lane_expect = mul(act, sel_val);
        result[ix++] = mul(mul(mul(p_map, pa), b_sponge), sub(cur[c.lanes_start + lane], lane_expect));
      }
    }
  }

  // shared gates
  F s_low = mul(p_last, p_map);                       // mixers.rs:19-25
  F g_carry = add(p_map, sub(p_pad, p_pad_last));     // ctrl.rs:370-373 / alu.rs:150-153
  for (int j = 0; j < ROUNDS - 1; j++) g_carry = add(g_carry, per[P_ROUND0 + j]);

  if (air.f_vm) {
    // ---------------- VmCtrlAir (ctrl.rs:114-390)
    F pi_prog = cur[c.pi_prog];
    F s_high = mul(s_low, pi_prog);
    F sum_dst0 = zero, sum_a = zero, sum_b = zero, sum_c = zero, sum_dst1 = zero;
    for (int i = 0; i < ZKL_NR; i++) {
      F sd0 = cur[c.sel_dst0_start + i], sa = cur[c.sel_a_start + i], sb = cur[c.sel_b_start + i],
        sc = cur[c.sel_c_start + i], sd1 = cur[c.sel_dst1_start + i];
      sum_dst0 = add(sum_dst0, sd0); sum_a = add(sum_a, sa); sum_b = add(sum_b, sb);
      sum_c = add(sum_c, sc); sum_dst1 = add(sum_dst1, sd1);
      result[ix++] = add(mul(mul(p_map, sd0), sub(sd0, one)), s_high);
      result[ix++] = add(mul(mul(p_map, sa), sub(sa, one)), s_high);
      result[ix++] = add(mul(mul(p_map, sb), sub(sb, one)), s_high);
      result[ix++] = add(mul(mul(p_map, sc), sub(sc, one)), s_high);
      result[ix++] = add(mul(mul(p_map, sd1), sub(sd1, one)), s_high);
    }
    F uses_a = zero, uses_b = zero, uses_c = zero, op_any = zero;
    for (int k : {MOV, ADD, SUB, MUL, NEG, EQ, SEL, DIVMOD, DIV128, MULWIDE, LOAD, STORE}) uses_a = add(uses_a, OP(k));
    for (int k : {ADD, SUB, MUL, EQ, SEL, DIVMOD, DIV128, MULWIDE, STORE}) uses_b = add(uses_b, OP(k));
    for (int k : {SEL, ASSERT, ASSERT_BIT, ASSERT_RANGE}) uses_c = add(uses_c, OP(k));
    for (int k : {CONST_, MOV, ADD, SUB, MUL, NEG, EQ, SEL, SPONGE, ASSERT, ASSERT_BIT, ASSERT_RANGE, DIVMOD, DIV128, MULWIDE})
      op_any = add(op_any, OP(k));
    F uses_dst0 = add(sub(op_any, OP(SPONGE)), OP(LOAD));
    F uses_dst1 = add(add(OP(DIVMOD), OP(DIV128)), OP(MULWIDE));
    result[ix++] = add(mul(p_map, sub(sum_dst0, uses_dst0)), s_low);
    result[ix++] = add(mul(p_map, sub(sum_a, uses_a)), s_low);
    result[ix++] = add(mul(p_map, sub(sum_b, uses_b)), s_low);
    result[ix++] = add(mul(p_map, sub(sum_c, uses_c)), s_low);
    result[ix++] = add(mul(p_map, sub(sum_dst1, uses_dst1)), s_low);
    for (int i = 0; i < ZKL_NR; i++)
      result[ix++] = add(mul(mul(p_map, cur[c.sel_dst0_start + i]), cur[c.sel_dst1_start + i]), s_high);
    if (air.f_sponge && sponge_used) {
      for (int lane = 0; lane < 10; lane++) {
        for (int b = 0; b < 3; b++) {
          F bv = cur[c.sel_s_bits_start + lane * 3 + b];
          result[ix++] = add(mul(mul(p_map, bv), sub(bv, one)), s_high);
        }
        F act = cur[c.sel_s_active_start + lane];
        result[ix++] = add(mul(mul(p_map, act), sub(act, one)), s_high);
      }
    }
    result[ix++] = s_high;  // ctrl.rs:292-293
    F op_sum = zero;
    for (int k = 0; k < ZKL_NUM_OPS; k++) {
      F b = OP(k); op_sum = add(op_sum, b);
      result[ix++] = add(mul(mul(p_map, b), sub(b, one)), s_high);
    }
    result[ix++] = add(mul(mul(p_map, op_sum), sub(op_sum, one)), s_high);
    F rom_en = air.rom_enabled ? one : zero;
    for (int k = 0; k < ZKL_NUM_OPS; k++) {
      // ctrl.rs:341-363; when the layout has no rom_op block the gate is zero anyway
      F rom_b = air.rom_enabled ? cur[c.rom_op_start + k] : zero;
      result[ix++] = add(mul(mul(rom_en, p_map), sub(OP(k), rom_b)), s_high);
    }
    F pc_cur = cur[c.pc], pc_next = next[c.pc];
    result[ix++] = add(mul(rom_en, mul(g_carry, sub(pc_next, pc_cur))), s_low);
    result[ix++] = add(mul(rom_en, mul(p_pad_last, sub(pc_next, add(pc_cur, one)))), s_low);

    // ---------------- VmAluAir (alu.rs:108-354)
    uint32_t m = air.vm_usage_mask;
    bool use_assert = m & 1, use_assert_bit = m >> 1 & 1, use_assert_range = m >> 2 & 1, use_divmod = m >> 3 & 1,
         use_mulwide = m >> 4 & 1, use_div128 = m >> 5 & 1, use_eq = m >> 6 & 1;
    F pi = cur[c.pi_prog];
    F pi2 = sq(pi), pi4 = sq(pi2), pi6 = mul(pi4, pi2);
    F s_write = mul(s_low, pi6), s_eq = mul(s_low, pi4);
    F a_val = zero, b_val = zero, c_val = zero;
    for (int i = 0; i < ZKL_NR; i++) {
      F r = cur[c.r_start + i];
      a_val = add(a_val, mul(cur[c.sel_a_start + i], r));
      b_val = add(b_val, mul(cur[c.sel_b_start + i], r));
      c_val = add(c_val, mul(cur[c.sel_c_start + i], r));
    }
    for (int i = 0; i < ZKL_NR; i++)
      result[ix++] = add(mul(g_carry, sub(next[c.r_start + i], cur[c.r_start + i])), s_low);
    F imm = cur[c.imm];
    F mode64 = cur[c.eq_inv];
    F dst0_next = zero, dst0_cur = zero, dst1_next = zero;
    for (int i = 0; i < ZKL_NR; i++) {
      dst0_next = add(dst0_next, mul(cur[c.sel_dst0_start + i], next[c.r_start + i]));
      dst0_cur = add(dst0_cur, mul(cur[c.sel_dst0_start + i], cur[c.r_start + i]));
      dst1_next = add(dst1_next, mul(cur[c.sel_dst1_start + i], next[c.r_start + i]));
    }
    F res = mul(OP(CONST_), imm);
    res = add(res, mul(OP(MOV), a_val));
    res = add(res, mul(OP(ADD), add(a_val, b_val)));
    res = add(res, mul(OP(SUB), sub(a_val, b_val)));
    res = add(res, mul(OP(MUL), mul(a_val, b_val)));
    res = add(res, mul(OP(NEG), sub(zero, a_val)));
    res = add(res, mul(OP(SEL), add(mul(c_val, a_val), mul(sub(one, c_val), b_val))));
    res = add(res, mul(OP(SPONGE), cur[c.lanes_start + 0]));
    if (use_eq) res = add(res, mul(OP(EQ), dst0_next));
    if (use_assert) res = add(res, OP(ASSERT));
    if (use_assert_bit) res = add(res, OP(ASSERT_BIT));
    res = add(res, mul(OP(LOAD), imm));
    F sum = zero, pow2 = one;
    for (int i = 0; i < 32; i++) { sum = add(sum, mul(pow2, cur[c.gadget_b_start + i])); pow2 = add(pow2, pow2); }
    if (use_assert_range) res = add(res, mul(OP(ASSERT_RANGE), add(mul(sub(one, imm), sum), imm)));
    bool uses_two = use_divmod || use_mulwide || use_div128;
    F b_two = uses_two ? add(add(OP(DIVMOD), OP(MULWIDE)), OP(DIV128)) : zero;
    for (int i = 0; i < ZKL_NR; i++) {
      F sd0 = cur[c.sel_dst0_start + i], sd1 = cur[c.sel_dst1_start + i];
      F keep = sub(sub(one, sd0), sd1);
      F w0 = add(mul(sub(one, b_two), res), mul(b_two, dst0_next));
      F w1 = mul(b_two, dst1_next);
      F rhs = add(add(mul(keep, cur[c.r_start + i]), mul(sd0, w0)), mul(sd1, w1));
      result[ix++] = add(mul(p_final, sub(next[c.r_start + i], rhs)), s_write);
    }
    F diff = sub(a_val, b_val), inv = cur[c.eq_inv];
    if (use_eq) {
      result[ix++] = add(mul(mul(p_final, OP(EQ)), mul(dst0_next, diff)), s_eq);
      result[ix++] = add(mul(mul(p_final, OP(EQ)), sub(sub(one, dst0_next), mul(diff, inv))), s_eq);
    }
    if (use_divmod) {
      result[ix++] = add(mul(mul(p_final, OP(DIVMOD)), sub(sub(a_val, mul(b_val, dst0_next)), dst1_next)), s_eq);
      result[ix++] = add(mul(mul(p_final, OP(DIVMOD)), sub(mul(b_val, inv), one)), s_eq);
    }
    F p2_64 = pow2_64();
    if (use_mulwide)
      result[ix++] = add(mul(mul(p_final, OP(MULWIDE)), sub(mul(a_val, b_val), add(dst0_next, mul(dst1_next, p2_64)))), s_eq);
    if (use_div128) {
      F num128 = add(mul(a_val, p2_64), imm);
      result[ix++] = add(mul(mul(p_final, OP(DIV128)), sub(num128, add(mul(b_val, dst0_next), dst1_next))), s_eq);
      result[ix++] = add(mul(mul(p_final, OP(DIV128)), sub(mul(b_val, inv), one)), s_eq);
    }
    if (use_assert)
      result[ix++] = add(mul(p_final, add(mul(OP(ASSERT), sub(c_val, one)), mul(OP(SEL), mul(c_val, sub(c_val, one))))), s_eq);
    if (use_assert_bit) result[ix++] = add(mul(mul(p_final, OP(ASSERT_BIT)), mul(c_val, sub(c_val, one))), s_eq);
    if (use_assert_range) {
      for (int i = 0; i < 32; i++) {
        F bi = cur[c.gadget_b_start + i];
        result[ix++] = add(mul(mul(p_final, OP(ASSERT_RANGE)), mul(bi, sub(bi, one))), s_eq);
      }
      F p2_32 = f_from_u64(1ULL << 32);
      F eq32 = sub(c_val, sum), eq64 = sub(c_val, add(dst0_cur, mul(sum, p2_32)));
      F eq_term = mul(imm, add(mul(mode64, eq64), mul(sub(one, mode64), eq32)));
      result[ix++] = add(mul(mul(p_final, OP(ASSERT_RANGE)), eq_term), s_eq);
    }
  }

  // ---------------- RamAir (ram.rs:82-236)
  if (air.f_ram) {
    const uint32_t R = c.ram_start;  // sorted, s_addr, s_clk, s_val, s_is_write, s_last_write, gp_unsorted, gp_sorted
    F g_hold = sub(p_pad, p_pad_last);
    F op_load = OP(LOAD), op_store = OP(STORE);
    F event = mul(p_final, add(op_load, op_store));
    F pi0 = FE(air.program_fe0);
    F pi2 = sq(pi0), pi3 = mul(pi2, pi0), pi4 = sq(pi2), pi5 = mul(pi4, pi0);
    F r1 = add(pi2, one), r2 = add(pi3, pi0), r3 = add(pi5, f_from_u64(7));
    F a_ev = zero, b_ev = zero;
    for (int i = 0; i < ZKL_NR; i++) {
      F ri = cur[c.r_start + i];
      a_ev = add(a_ev, mul(cur[c.sel_a_start + i], ri));
      b_ev = add(b_ev, mul(cur[c.sel_b_start + i], ri));
    }
    F w_ev = op_store;
    F val_ev = add(mul(w_ev, b_ev), mul(sub(one, w_ev), cur[c.imm]));
    F clk_ev = cur[c.pc];
    F comp_uns = add(add(add(a_ev, mul(r1, clk_ev)), mul(r2, val_ev)), mul(r3, w_ev));
    F gu = cur[R + 6], gu_n = next[R + 6], gs = cur[R + 7], gs_n = next[R + 7];
    result[ix++] = add(add(mul(event, sub(gu_n, add(gu, comp_uns))), mul(sub(one, event), sub(gu_n, gu))), mul(g_hold, sub(gu_n, gu)));
    F s_on = cur[R + 0], s_addr = cur[R + 1], s_clk = cur[R + 2], s_val = cur[R + 3], s_w = cur[R + 4], last = cur[R + 5];
    F s_addr_n = next[R + 1], s_clk_n = next[R + 2], last_n = next[R + 5];
    F d_addr = sub(s_addr_n, s_addr);
    F inv = cur[c.eq_inv];
    F same = sub(one, mul(d_addr, inv));
    F comp = add(add(add(s_addr, mul(r1, s_clk)), mul(r2, s_val)), mul(r3, s_w));
    result[ix++] = add(mul(s_on, sub(gs_n, add(gs, comp))), mul(sub(one, s_on), sub(gs_n, gs)));
    F last_keep = add(mul(same, add(mul(sub(one, s_w), last), mul(s_w, s_val))), mul(sub(one, same), mul(s_w, s_val)));
    result[ix++] = mul(s_on, sub(last_n, last_keep));
    result[ix++] = mul(mul(s_on, sub(one, s_w)), sub(s_val, last));
    F s_on_n = next[R + 0], s_w_n = next[R + 4], s_val_n = next[R + 3];
    result[ix++] = mul(mul(mul(mul(s_on, s_on_n), sub(one, same)), sub(one, s_w_n)), s_val_n);
    result[ix++] = mul(s_on, mul(same, sub(same, one)));
    if ((air.vm_usage_mask >> 8) & 1) {
      F d_clk = sub(s_clk_n, s_clk);
      F sum = zero, pow2 = one;
      for (int i = 0; i < 32; i++) {
        F bi = cur[c.gadget_b_start + i];
        if ((air.ram_delta_clk_bits >> i) & 1) result[ix++] = mul(mul(mul(s_on, same), bi), sub(bi, one));
        sum = add(sum, mul(pow2, bi)); pow2 = add(pow2, pow2);
      }
      result[ix++] = mul(mul(mul(s_on, s_on_n), same), sub(d_clk, sum));
    }
    result[ix++] = mul(p_last, sub(gu, gs));
  }

  // ---------------- MerkleAir (merkle.rs:66-134)
  if (air.f_merkle) {
    const uint32_t M = c.merkle_start;  // g, dir, sib, acc, first, last, leaf
    F g = cur[M + 0], dir = cur[M + 1], sib = cur[M + 2], acc = cur[M + 3];
    result[ix++] = mul(mul(mul(p_map, g), dir), sub(dir, one));
    F left = add(mul(sub(one, dir), acc), mul(dir, sib));
    F right = add(mul(sub(one, dir), sib), mul(dir, acc));
    result[ix++] = mul(mul(p_map, g), sub(cur[c.lanes_start + 0], left));
    result[ix++] = mul(mul(p_map, g), sub(cur[c.lanes_start + 1], right));
    result[ix++] = mul(mul(g, g_carry), sub(next[M + 3], acc));   // g_hold == g_carry (merkle.rs:103-106)
    result[ix++] = mul(mul(mul(p_map, g), cur[M + 4]), sub(acc, cur[M + 6]));
    result[ix++] = mul(mul(mul(p_final, g), cur[M + 5]), sub(acc, FE(air.merkle_root_fe)));
    result[ix++] = mul(mul(mul(p_pad_last, g), next[M + 0]), sub(next[M + 3], acc));
  }

  // ---------------- RomAir (rom.rs:57-120); gate = program_commitment != 0 (mod.rs:355)
  if (air.rom_enabled) {
    F s3[3];
    for (int i = 0; i < 3; i++) s3[i] = f_cube(cur[c.rom_s_start + i]);
    for (int j = 0; j < ROUNDS; j++) {
      F gr = per[P_ROUND0 + j];
      for (int i = 0; i < 3; i++) {
        F acc = zero;
        for (int k = 0; k < 3; k++) acc = add(acc, mul(FE(air.rom.mds[i][k]), s3[k]));
        F y = add(acc, FE(air.rom.rc[j][i]));
        result[ix++] = mul(gr, sub(next[c.rom_s_start + i], y));
      }
    }
    F g_hold = sub(p_pad, p_pad_last);
    for (int i = 0; i < 3; i++) result[ix++] = mul(g_hold, sub(next[c.rom_s_start + i], cur[c.rom_s_start + i]));
    // rom.rs:104-119: both arms equal p_map*(...) (zero when p_map == 0)
    F enc0 = zero, enc1 = zero;
    int k = 0;
    auto acc_enc = [&](uint32_t col) {
      enc0 = add(enc0, mul(cur[col], FE(air.rom_w_enc0[k])));
      enc1 = add(enc1, mul(cur[col], FE(air.rom_w_enc1[k])));
      k++;
    };
    for (int o = 0; o < ZKL_NUM_OPS; o++) acc_enc(c.op_start + o);
    for (int i = 0; i < ZKL_NR; i++) acc_enc(c.sel_dst0_start + i);
    for (int i = 0; i < ZKL_NR; i++) acc_enc(c.sel_a_start + i);
    for (int i = 0; i < ZKL_NR; i++) acc_enc(c.sel_b_start + i);
    for (int i = 0; i < ZKL_NR; i++) acc_enc(c.sel_c_start + i);
    for (int i = 0; i < ZKL_NR; i++) acc_enc(c.sel_dst1_start + i);
    result[ix++] = mul(p_map, sub(cur[c.rom_s_start + 1], enc0));
    result[ix++] = mul(p_map, sub(cur[c.rom_s_start + 2], enc1));
  }
  // vm/air/mod.rs:359-377: count must match the declared degree list
  if (ix != air.num_transition) { fprintf(stderr, "oracle: constraint count mismatch %zu vs %u\n", ix, air.num_transition); abort(); }
}

}  // namespace zkl
