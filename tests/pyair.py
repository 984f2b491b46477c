"""Independent Python-integer restatement of four blocks of ZkLispAir::evaluate_transition, written from the reference text
(not from oracle/air.h): PoseidonAir (vm/air/poseidon.rs:67-162), RamAir (ram.rs:82-236), MerkleAir (merkle.rs:66-134), RomAir
(rom.rs:57-120) with rom_linear_encode_from_slice (utils.rs:144-208). It pins the checker's constraint ORDER and formulas: an
AIR-valid trace makes every constraint zero whatever their order, so only an evaluation on random frames compared index by
index can catch a swapped or mis-weighted constraint. VmCtrlAir / VmAluAir are not restated here."""
from zk_lisp_b200.air import NR, POSEIDON_ROUNDS
from zk_lisp_b200.field import P

R = POSEIDON_ROUNDS


def poseidon_block(a, cur, nxt, per):
    c, s = a.cols, a.pose
    out = []
    p_map = per[0]
    pa = cur[c.pose_active]
    st = [cur[c.lane_index(i)] for i in range(12)]
    s3 = [pow(v, 3, P) for v in st]
    for j in range(R):
        gr = per[1 + j]
        for i in range(12):
            y = (sum(s.mds[i][k] * s3[k] for k in range(12)) + s.rc[j][i]) % P
            out.append(pa * gr * (nxt[c.lane_index(i)] - y) % P)
    g_hold = (per[1 + R + 1] - per[1 + R + 2]) % P
    for i in range(12):
        out.append(g_hold * (nxt[c.lane_index(i)] - cur[c.lane_index(i)]) % P)
    sponge_used = (a.air_pi.vm_usage_mask >> 7) & 1
    if a.fm & 2 and a.fm & 32 and sponge_used:          # FM_VM, FM_SPONGE
        b_sponge = cur[c.op_sponge]
        r = [cur[c.r_index(i)] for i in range(8)]
        for lane in range(10):
            b0, b1, b2 = (cur[c.sel_s_bits_start + 3 * lane + k] for k in range(3))
            act = cur[c.sel_s_active_start + lane]
            s0, s1 = b0 * r[1] + (1 - b0) * r[0], b0 * r[3] + (1 - b0) * r[2]
            s2, s3_ = b0 * r[5] + (1 - b0) * r[4], b0 * r[7] + (1 - b0) * r[6]
            t0, t1 = b1 * s1 + (1 - b1) * s0, b1 * s3_ + (1 - b1) * s2
            sel_val = b2 * t1 + (1 - b2) * t0
            out.append(p_map * pa * b_sponge * (cur[c.lane_index(lane)] - act * sel_val) % P)
    return out


def ram_block(a, cur, nxt, per):
    c = a.cols
    rs = c.ram_start      # sorted, s_addr, s_clk, s_val, s_is_write, s_last_write, gp_unsorted, gp_sorted
    p_final, p_pad, p_pad_last, p_last = per[1 + R], per[1 + R + 1], per[1 + R + 2], per[1 + R + 3]
    g_hold = p_pad - p_pad_last
    op_load, op_store = cur[c.op_load], cur[c.op_store]
    event = p_final * (op_load + op_store)
    pi0 = a.program_fe[0]
    r1, r2, r3 = pi0 ** 2 + 1, pi0 ** 3 + pi0, pi0 ** 5 + 7
    a_ev = sum(cur[c.sel_a_start + i] * cur[c.r_index(i)] for i in range(NR))
    b_ev = sum(cur[c.sel_b_start + i] * cur[c.r_index(i)] for i in range(NR))
    w_ev = op_store
    val_ev = w_ev * b_ev + (1 - w_ev) * cur[c.imm]
    comp_uns = a_ev + r1 * cur[c.pc] + r2 * val_ev + r3 * w_ev
    gu, gs = rs + 6, rs + 7
    out = [(event * (nxt[gu] - (cur[gu] + comp_uns)) + (1 - event) * (nxt[gu] - cur[gu]) + g_hold * (nxt[gu] - cur[gu])) % P]
    s_on, s_addr, s_clk, s_val, s_w, last = (cur[rs + k] for k in range(6))
    same = 1 - (nxt[rs + 1] - s_addr) * cur[c.eq_inv]
    comp = s_addr + r1 * s_clk + r2 * s_val + r3 * s_w
    out.append((s_on * (nxt[gs] - (cur[gs] + comp)) + (1 - s_on) * (nxt[gs] - cur[gs])) % P)
    last_keep = same * ((1 - s_w) * last + s_w * s_val) + (1 - same) * (s_w * s_val)
    out.append(s_on * (nxt[rs + 5] - last_keep) % P)
    out.append(s_on * (1 - s_w) * (s_val - last) % P)
    out.append(s_on * nxt[rs] * (1 - same) * (1 - nxt[rs + 4]) * nxt[rs + 3] % P)
    out.append(s_on * (same * (same - 1)) % P)
    if (a.air_pi.vm_usage_mask >> 8) & 1:
        d_clk = nxt[rs + 2] - s_clk
        total = 0
        for i in range(32):
            bi = cur[c.gadget_b_start + i]
            if (a.air_pi.ram_delta_clk_bits >> i) & 1:
                out.append(s_on * same * bi * (bi - 1) % P)
            total += (1 << i) * bi
        out.append(s_on * nxt[rs] * same * (d_clk - total) % P)
    out.append(p_last * (cur[gu] - cur[gs]) % P)
    return out


def merkle_block(a, cur, nxt, per, merkle_root_fe):
    c = a.cols
    m = c.merkle_start    # g, dir, sib, acc, first, last, leaf
    p_map, p_final, p_pad, p_pad_last = per[0], per[1 + R], per[1 + R + 1], per[1 + R + 2]
    g, d, sib, acc = cur[m], cur[m + 1], cur[m + 2], cur[m + 3]
    out = [p_map * g * d * (d - 1) % P]
    left, right = (1 - d) * acc + d * sib, (1 - d) * sib + d * acc
    out.append(p_map * g * (cur[c.lane_l] - left) % P)
    out.append(p_map * g * (cur[c.lane_r] - right) % P)
    g_hold = p_map + (p_pad - p_pad_last) + sum(per[1 + j] for j in range(R - 1))
    out.append(g * g_hold * (nxt[m + 3] - acc) % P)
    out.append(p_map * g * cur[m + 4] * (acc - cur[m + 6]) % P)
    out.append(p_final * g * cur[m + 5] * (acc - merkle_root_fe) % P)
    out.append(p_pad_last * g * nxt[m] * (nxt[m + 3] - acc) % P)
    return out


def rom_encode(a, cur, weights):
    c = a.cols
    cols = [c.op_start + k for k in range(17)]
    for start in (c.sel_dst0_start, c.sel_a_start, c.sel_b_start, c.sel_c_start, c.sel_dst1_start):
        cols += [start + i for i in range(NR)]
    return sum(cur[col] * w for col, w in zip(cols, weights)) % P


def rom_block(a, cur, nxt, per, w0, w1):
    c = a.cols
    out = []
    s = [cur[c.rom_s_index(i)] for i in range(3)]
    s3 = [pow(v, 3, P) for v in s]
    for j in range(R):
        for i in range(3):
            y = (sum(a.rom_mds[i][k] * s3[k] for k in range(3)) + a.rom_rc[j][i]) % P
            out.append(per[1 + j] * (nxt[c.rom_s_index(i)] - y) % P)
    g_hold = per[1 + R + 1] - per[1 + R + 2]
    for i in range(3):
        out.append(g_hold * (nxt[c.rom_s_index(i)] - s[i]) % P)
    p_map = per[0]
    out.append(p_map * (s[1] - rom_encode(a, cur, w0)) % P)
    out.append(p_map * (s[2] - rom_encode(a, cur, w1)) % P)
    return out
