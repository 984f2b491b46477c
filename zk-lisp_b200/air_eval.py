"""ZkLispAir on Python integers: `evaluate_transition` (vm/air/mod.rs:324-357) with its six blocks — PoseidonAir
(vm/air/poseidon.rs:67-162), VmCtrlAir (ctrl.rs:114-390), VmAluAir (alu.rs:108-354), RamAir (ram.rs:82-236), MerkleAir
(merkle.rs:66-134), RomAir (rom.rs:57-120, rom_linear_encode_from_slice utils.rs:144-208), mixers (mixers.rs:19-48) — and
`get_assertions` (mod.rs:380-504, schedule.rs:42-140, rom.rs:122-147). Written from the reference text, independently of the
CUDA kernel (csrc/kernels_air.cuh) and of the test checker.

Two uses: the host verifier (verify.py) evaluates the constraints at ONE out-of-domain point per proof with it, and the test suite
pins the constraint ORDER and formulas of the other two implementations against it on random frames, index by index — an
AIR-valid trace makes every constraint zero whatever their order, so only such a comparison catches a swapped constraint."""
from .air import NR, POSEIDON_ROUNDS
from .field import P

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


# ---- VmCtrlAir (ctrl.rs:114-390) and VmAluAir (alu.rs:108-354) -----------------------------------------------------
OP_ORDER = ["const", "mov", "add", "sub", "mul", "neg", "eq", "select", "sponge", "assert", "assert_bit", "assert_range",
            "divmod", "div128", "mulwide", "load", "store"]          # the order of the op booleanity / ROM lists (ctrl.rs:262-340)


def _low(per):                       # mixers.rs:19-24
    return per[1 + R + 3] * per[0] % P


def ctrl_block(a, cur, nxt, per, rom_enabled):
    c = a.cols
    p_map = per[0]
    pi_prog = cur[c.pi_prog]
    s_low, s_high = _low(per), _low(per) * pi_prog % P
    b = {name: cur[getattr(c, "op_" + name)] for name in OP_ORDER}
    out = []
    sums = [0] * 5
    starts = (c.sel_dst0_start, c.sel_a_start, c.sel_b_start, c.sel_c_start, c.sel_dst1_start)
    for i in range(NR):
        for k, st in enumerate(starts):
            v = cur[st + i]
            sums[k] += v
            out.append((p_map * v * (v - 1) + s_high) % P)
    uses_a = sum(b[k] for k in ("mov", "add", "sub", "mul", "neg", "eq", "select", "divmod", "div128", "mulwide", "load", "store"))
    uses_b = sum(b[k] for k in ("add", "sub", "mul", "eq", "select", "divmod", "div128", "mulwide", "store"))
    uses_c = b["select"] + b["assert"] + b["assert_bit"] + b["assert_range"]
    op_any = sum(b[k] for k in OP_ORDER[:15])
    uses_dst0 = op_any - b["sponge"] + b["load"]
    uses_dst1 = b["divmod"] + b["div128"] + b["mulwide"]
    for total, uses in zip(sums, (uses_dst0, uses_a, uses_b, uses_c, uses_dst1)):
        out.append((p_map * (total - uses) + s_low) % P)
    for i in range(NR):
        out.append((p_map * cur[c.sel_dst0_start + i] * cur[c.sel_dst1_start + i] + s_high) % P)
    if a.fm & 32 and (a.air_pi.vm_usage_mask >> 7) & 1:       # FM_SPONGE and sponge used
        for lane in range(10):
            for k in range(3):
                v = cur[c.sel_s_bits_start + 3 * lane + k]
                out.append((p_map * v * (v - 1) + s_high) % P)
            act = cur[c.sel_s_active_start + lane]
            out.append((p_map * act * (act - 1) + s_high) % P)
    out.append(s_high)
    for name in OP_ORDER:
        out.append((p_map * b[name] * (b[name] - 1) + s_high) % P)
    op_sum = sum(b.values())
    out.append((p_map * op_sum * (op_sum - 1) + s_high) % P)
    for k, name in enumerate(OP_ORDER):
        out.append((rom_enabled * p_map * (b[name] - cur[c.rom_op_start + k]) + s_high) % P)
    g_carry = p_map + (per[1 + R + 1] - per[1 + R + 2]) + sum(per[1 + j] for j in range(R - 1))
    out.append((rom_enabled * (g_carry * (nxt[c.pc] - cur[c.pc])) + s_low) % P)
    out.append((rom_enabled * (per[1 + R + 2] * (nxt[c.pc] - (cur[c.pc] + 1))) + s_low) % P)
    return out


def alu_block(a, cur, nxt, per):
    c = a.cols
    mask = a.air_pi.vm_usage_mask
    use = lambda bit: (mask >> bit) & 1               # layout.rs VM_USAGE_*: assert 0, assert_bit 1, assert_range 2, divmod 3,
    use_assert, use_assert_bit, use_assert_range = use(0), use(1), use(2)        # mulwide 4, div128 5, eq 6
    use_divmod, use_mulwide, use_div128, use_eq = use(3), use(4), use(5), use(6)
    p_map, p_final, p_pad, p_pad_last = per[0], per[1 + R], per[1 + R + 1], per[1 + R + 2]
    pi = cur[c.pi_prog]
    s_low = _low(per)
    s_write, s_eq = s_low * pow(pi, 6, P) % P, s_low * pow(pi, 4, P) % P
    g_carry = p_map + (p_pad - p_pad_last) + sum(per[1 + j] for j in range(R - 1))
    regs = [cur[c.r_index(i)] for i in range(NR)]
    a_val = sum(cur[c.sel_a_start + i] * regs[i] for i in range(NR))
    b_val = sum(cur[c.sel_b_start + i] * regs[i] for i in range(NR))
    c_val = sum(cur[c.sel_c_start + i] * regs[i] for i in range(NR))
    out = [(g_carry * (nxt[c.r_index(i)] - regs[i]) + s_low) % P for i in range(NR)]
    imm = cur[c.imm]
    b = {name: cur[getattr(c, "op_" + name)] for name in OP_ORDER}
    mode64 = cur[c.eq_inv]
    dst0_next = sum(cur[c.sel_dst0_start + i] * nxt[c.r_index(i)] for i in range(NR))
    dst0_cur = sum(cur[c.sel_dst0_start + i] * regs[i] for i in range(NR))
    dst1_next = sum(cur[c.sel_dst1_start + i] * nxt[c.r_index(i)] for i in range(NR))
    res = (b["const"] * imm + b["mov"] * a_val + b["add"] * (a_val + b_val) + b["sub"] * (a_val - b_val) + b["mul"] * (a_val * b_val)
           + b["neg"] * (-a_val) + b["select"] * (c_val * a_val + (1 - c_val) * b_val) + b["sponge"] * cur[c.lane_l]
           + (b["eq"] * dst0_next if use_eq else 0) + (b["assert"] if use_assert else 0) + (b["assert_bit"] if use_assert_bit else 0)
           + b["load"] * imm)
    bits_sum = sum((1 << i) * cur[c.gadget_b_start + i] for i in range(32))
    if use_assert_range:
        res += b["assert_range"] * ((1 - imm) * bits_sum + imm)
    b_two = (b["divmod"] + b["mulwide"] + b["div128"]) if (use_divmod or use_mulwide or use_div128) else 0
    for i in range(NR):
        sd0, sd1 = cur[c.sel_dst0_start + i], cur[c.sel_dst1_start + i]
        keep = 1 - sd0 - sd1
        w0 = (1 - b_two) * res + b_two * dst0_next
        w1 = b_two * dst1_next
        out.append((p_final * (nxt[c.r_index(i)] - (keep * regs[i] + sd0 * w0 + sd1 * w1)) + s_write) % P)
    diff, inv_ = a_val - b_val, cur[c.eq_inv]
    if use_eq:
        out.append((p_final * b["eq"] * (dst0_next * diff) + s_eq) % P)
        out.append((p_final * b["eq"] * ((1 - dst0_next) - diff * inv_) + s_eq) % P)
    if use_divmod:
        out.append((p_final * b["divmod"] * (a_val - b_val * dst0_next - dst1_next) + s_eq) % P)
        out.append((p_final * b["divmod"] * (b_val * inv_ - 1) + s_eq) % P)
    p2_64 = 1 << 64
    if use_mulwide:
        out.append((p_final * b["mulwide"] * (a_val * b_val - (dst0_next + dst1_next * p2_64)) + s_eq) % P)
    if use_div128:
        out.append((p_final * b["div128"] * ((a_val * p2_64 + imm) - (b_val * dst0_next + dst1_next)) + s_eq) % P)
        out.append((p_final * b["div128"] * (b_val * inv_ - 1) + s_eq) % P)
    if use_assert:
        out.append((p_final * (b["assert"] * (c_val - 1) + b["select"] * (c_val * (c_val - 1))) + s_eq) % P)
    if use_assert_bit:
        out.append((p_final * b["assert_bit"] * (c_val * (c_val - 1)) + s_eq) % P)
    if use_assert_range:
        for i in range(32):
            bi = cur[c.gadget_b_start + i]
            out.append((p_final * b["assert_range"] * (bi * (bi - 1)) + s_eq) % P)
        eq32 = c_val - bits_sum
        eq64 = c_val - (dst0_cur + bits_sum * (1 << 32))
        out.append((p_final * b["assert_range"] * (imm * (mode64 * eq64 + (1 - mode64) * eq32)) + s_eq) % P)
    return out


def evaluate_transition(a, cur, nxt, per, w0, w1, merkle_root_fe, rom_enabled=1):
    """vm/air/mod.rs:324-357: the whole vector in block order."""
    out = []
    if a.fm & 1:
        out += poseidon_block(a, cur, nxt, per)
    if a.fm & 2:
        out += ctrl_block(a, cur, nxt, per, rom_enabled) + alu_block(a, cur, nxt, per)
    if a.fm & 128:
        out += ram_block(a, cur, nxt, per)
    if a.fm & 64:
        out += merkle_block(a, cur, nxt, per, merkle_root_fe)
    if rom_enabled:
        out += rom_block(a, cur, nxt, per, w0, w1)
    return out


def assertions(cols, n, air_pi, fm):
    """get_assertions (vm/air/mod.rs:380-504) with ScheduleAir (schedule.rs:42-140) and RomAir (rom.rs:122-147), restated from the
    reference text: (column, row, value) in emission order, first occurrence of a (column, row) kept."""
    from . import params
    from .air import encode_main_args_to_slots
    from .field import be_from_le8
    core = air_pi.core
    last = n - 1
    dom = params.get_poseidon_suite(bytes(core.program_id)).dom
    rom = any(core.program_commitment)
    out = []
    for lvl in range((last + 1) // 32):
        base = lvl * 32
        row_map, row_final = base, base + 1 + R
        out += [(cols.lane_c0, row_map, dom[0]), (cols.lane_c1, row_map, dom[1]), (cols.g_map, row_map, 1), (cols.g_final, row_final, 1)]
        out += [(cols.g_r_index(j), base + 1 + j, 1) for j in range(R)]
        out.append((cols.g_final, row_map, 0))
        out += [(cols.g_r_index(j), row_map, 0) for j in range(R)]
        out.append((cols.g_map, row_final, 0))
        out += [(cols.g_r_index(j), row_final, 0) for j in range(R)]
        for j in range(R):
            out += [(cols.g_map, base + 1 + j, 0), (cols.g_final, base + 1 + j, 0)]
        if lvl == 0 and fm & 2:
            if air_pi.pc_init == 0 and rom:
                out.append((cols.pi_prog, row_map, be_from_le8(core.program_commitment)))
            out.append((cols.pc, row_map, air_pi.pc_init))
    if fm & 2:
        if fm & 16:      # FM_VM_EXPECT
            out.append((cols.r_index(min(core.vm_out_reg, NR - 1)), min(core.vm_out_row, last), be_from_le8(core.vm_expected_bytes)))
        slots = encode_main_args_to_slots(core.main_args)
        if air_pi.pc_init == 0 and slots:
            assert len(slots) <= NR
            out += [(cols.r_index(NR - len(slots) + j), 0, v) for j, v in enumerate(slots)]
    if rom:
        out += [(cols.rom_s_index(i), 0, air_pi.rom_s_in[i]) for i in range(3)]
        out += [(cols.rom_s_index(i), last, air_pi.rom_s_out[i]) for i in range(3)]
    if not out:
        out.append((cols.mask, last, 0))
    seen, dedup = set(), []
    for col, row, val in out:
        if (col, row) not in seen:
            seen.add((col, row))
            dedup.append((col, row, val % P))
    return dedup
