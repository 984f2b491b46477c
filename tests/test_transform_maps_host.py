"""CPU models (Python integers) of the index maps and constant tables the tensor-core transform kernels implement
(zk-lisp_b200/csrc/kernels_tcntt.cuh) and of the DEEP stage's batch inversion (kernels_stark.cuh). The kernels themselves
are checked on the GPU against the oracle; these tests pin the ALGEBRA — which twiddle goes where, that the coset shift
can live in the per-coset matrix B_k, that using B_k in all three stages is undone by the factors folded into tw0 —
independently of any device code, so a table-builder change that breaks it fails here, without a GPU."""
import random

from zk_lisp_b200.field import P, inv, root_of_unity

GENERATOR = 3


def _dft(a, w):
    """Naive O(n log n) recursive DFT: X[i] = sum_j a[j] w^(i j)."""
    n = len(a)
    if n == 1:
        return a[:]
    ev, od = _dft(a[0::2], w * w % P), _dft(a[1::2], w * w % P)
    out, t = [0] * n, 1
    for i in range(n // 2):
        x = t * od[i] % P
        out[i], out[i + n // 2] = (ev[i] + x) % P, (ev[i] - x) % P
        t = t * w % P
    return out


def _lde4096_v2(a, s, w4096):
    """k_tc_lde4096_v2 for one coset: y[i] = sum_j a_j (s w^i)^j, i < 4096, computed the way the kernel does.
    j = jA + 16 jC + 256 jD; the SAME matrix B[iX][jX] = (s^256)^jX w16^(iX jX) multiplies all three stages."""
    w16, w256 = pow(w4096, 256, P), pow(w4096, 16, P)
    s256 = pow(s, 256, P)
    B = [[pow(s256, j, P) * pow(w16, i * j, P) % P for j in range(16)] for i in range(16)]
    s256_inv = inv(s256)
    # k_build_lde16_consts: tw0[iD][m1] = s^m1 * w_256^(iD jC) / (s^256)^(jA + jC)
    tw0 = [[pow(s, m1, P) * pow(w256, iD * (m1 >> 4), P) % P * pow(s256_inv, (m1 & 15) + (m1 >> 4), P) % P for m1 in range(256)]
           for iD in range(16)]
    # stage 1: instance m1 = jA + 16 jC, inputs over jD -> chunk jC, instance jA + 16 iD
    st1 = [[0] * 256 for _ in range(16)]
    for m1 in range(256):
        jA, jC = m1 & 15, m1 >> 4
        for iD in range(16):
            acc = sum(a[m1 + 256 * jD] * B[iD][jD] for jD in range(16)) % P
            st1[jC][jA + 16 * iD] = acc * tw0[iD][m1] % P
    # stage 2: instance m2 = jA + 16 iD, inputs over jC (chunk), twiddle w^((16 iC + iD) jA) -> chunk jA, instance iB = 16 iC + iD
    st2 = [[0] * 256 for _ in range(16)]
    for m2 in range(256):
        jA, iD = m2 & 15, m2 >> 4
        for iC in range(16):
            iB = 16 * iC + iD
            acc = sum(st1[jC][m2] * B[iC][jC] for jC in range(16)) % P
            st2[jA][iB] = acc * pow(w4096, iB * jA, P) % P
    # stage 3: instance iB, inputs over jA -> y[256 iA + iB]
    y = [0] * 4096
    for iB in range(256):
        for iA in range(16):
            y[256 * iA + iB] = sum(st2[jA][iB] * B[iA][jA] for jA in range(16)) % P
    return y


def test_radix16_stages_with_the_coset_shift_folded_into_the_matrix():
    rnd = random.Random(1)
    w = root_of_unity(12 + 4)                       # w_N, N = 4096 * 16
    w4096 = pow(w, 16, P)
    a = [rnd.randrange(P) for _ in range(4096)]
    for k in (0, 5):
        s = GENERATOR * pow(w, k, P) % P            # coset k of the blowup-16 domain
        got = _lde4096_v2(a, s, w4096)
        shifted = [a[j] * pow(s, j, P) % P for j in range(4096)]
        assert got == _dft(shifted, w4096)


def test_inverse_variant_scales_once():
    """Interpolation: B = n^-1 * w16^-(ij) in all three stages, tw0 carries n^2: the result is scaled by n^-1 exactly once."""
    rnd = random.Random(2)
    w4096 = root_of_unity(12)
    wi = inv(w4096)
    ev = [rnd.randrange(P) for _ in range(4096)]
    n_inv = inv(4096)
    w16, w256 = pow(wi, 256, P), pow(wi, 16, P)
    B = [[n_inv * pow(w16, i * j, P) % P for j in range(16)] for i in range(16)]
    n2 = 4096 * 4096 % P
    st1 = [[0] * 256 for _ in range(16)]
    for m1 in range(256):
        for iD in range(16):
            acc = sum(ev[m1 + 256 * jD] * B[iD][jD] for jD in range(16)) % P
            st1[m1 >> 4][(m1 & 15) + 16 * iD] = acc * pow(w256, iD * (m1 >> 4), P) % P * n2 % P
    st2 = [[0] * 256 for _ in range(16)]
    for m2 in range(256):
        jA, iD = m2 & 15, m2 >> 4
        for iC in range(16):
            iB = 16 * iC + iD
            st2[jA][iB] = sum(st1[jC][m2] * B[iC][jC] for jC in range(16)) % P * pow(wi, iB * jA, P) % P
    coeffs = [0] * 4096
    for iB in range(256):
        for iA in range(16):
            coeffs[256 * iA + iB] = sum(st2[jA][iB] * B[iA][jA] for jA in range(16)) % P
    assert _dft(coeffs, w4096) == ev


def test_cross_stage_decomposition():
    """k_tc_cross16 + per-block coset LDE for n = R * m: y[i2 + R i1] = sum_j1 [s^j1 w_m^(j1 i1)] w_n^(j1 i2) sum_j2 a[j1 + m j2] (s^m)^j2 w_R^(j2 i2).
    Modelled at m = 64 (the structure does not depend on m) for R = 4, 8, 16."""
    rnd = random.Random(3)
    m = 64
    for R in (4, 8, 16):
        n = R * m
        logn = n.bit_length() - 1
        wN = root_of_unity(logn + 3)                 # blowup 8
        wn, s = pow(wN, 8, P), GENERATOR * pow(wN, 3, P) % P
        wm, wR = pow(wn, R, P), pow(wn, m, P)
        a = [rnd.randrange(P) for _ in range(n)]
        sm = pow(s, m, P)
        # pass 1: block i2 of the slab, element j1
        slab = [[sum(a[j1 + m * j2] * pow(sm, j2, P) * pow(wR, j2 * i2, P) for j2 in range(R)) % P * pow(wn, j1 * i2, P) % P
                 for j1 in range(m)] for i2 in range(R)]
        # pass 2: coset LDE (shift s) of every block, scattered with stride R
        y = [0] * n
        for i2 in range(R):
            blk = _dft([slab[i2][j1] * pow(s, j1, P) % P for j1 in range(m)], wm)
            for i1 in range(m):
                y[i2 + R * i1] = blk[i1]
        assert y == _dft([a[j] * pow(s, j, P) % P for j in range(n)], wn)


def test_batch_inversion_prefix_suffix():
    """k_deep_inverses: per-thread prefix products, prefix and suffix scans of the thread totals, ONE inversion per block."""
    rnd = random.Random(4)
    threads, per = 16, 8
    v = [[rnd.randrange(1, P) for _ in range(per)] for _ in range(threads)]
    p = [[0] * per for _ in range(threads)]
    for t in range(threads):
        acc = 1
        for j in range(per):
            acc = acc * v[t][j] % P
            p[t][j] = acc
    tot = [p[t][per - 1] for t in range(threads)]
    pre, suf = tot[:], tot[:]
    for t in range(1, threads):
        pre[t] = pre[t - 1] * tot[t] % P
    for t in range(threads - 2, -1, -1):
        suf[t] = suf[t + 1] * tot[t] % P
    block_inv = inv(pre[threads - 1])
    for t in range(threads):
        r = block_inv
        if t > 0:
            r = r * pre[t - 1] % P
        if t < threads - 1:
            r = r * suf[t + 1] % P
        for j in range(per - 1, -1, -1):
            iv = r * p[t][j - 1] % P if j else r
            if j:
                r = r * v[t][j] % P
            assert iv * v[t][j] % P == 1


def _multilevel(a, s, wn, radices, m):
    """dev_tc_multilevel for one coset, modelled on Python integers: the column of n = R_1 ... R_L m coefficients goes through L cross
    stages (per level: matrix entry (s^sub)^j2 w_R^(i2 j2), twiddle w_len^(j1 i2), block-major destination = the columns of the level
    below) and the m-point coset transform of every block; block b's outputs land at y[o(b) + P i1] with o the digit reversal
    k_tc_lde4096_v2 (scatter) and k_tc_unscatter (tiled transpose) both implement. Returns (y, the o(b) map)."""
    n = len(a)
    cols, length = [a[:]], n              # level l: `cols` columns of `length` elements
    w_len = wn
    for R in radices:
        sub = length // R
        s_sub, wR = pow(s, sub, P), pow(w_len, sub, P)
        nxt = []
        for col in cols:                  # column c -> columns c R + i2 of the level below
            for i2 in range(R):
                nxt.append([sum(col[j1 + sub * j2] * pow(s_sub, j2, P) * pow(wR, j2 * i2, P) for j2 in range(R)) % P * pow(w_len, j1 * i2, P) % P
                            for j1 in range(sub)])
        cols, length, w_len = nxt, sub, pow(w_len, R, P)
    assert length == m
    Pblocks = len(cols)
    y, omap = [0] * n, []
    for b, col in enumerate(cols):
        # o(b): b = ((i2_1 R_2 + i2_2) R_3 ...) + i2_L, o = i2_1 + R_1 (i2_2 + R_2 (...))
        digits, rem = [], b
        for R in reversed(radices):
            digits.append(rem % R); rem //= R
        digits.reverse()                  # i2_1 .. i2_L
        o, weight = 0, 1
        for d, R in zip(digits, radices):
            o += d * weight; weight *= R
        omap.append(o)
        blk = _dft([col[j1] * pow(s, j1, P) % P for j1 in range(m)], w_len)
        for i1 in range(m):
            y[o + Pblocks * i1] = blk[i1]
    return y, omap


def test_multilevel_cross_stages_and_the_digit_reversed_scatter():
    """n = R_1 ... R_L m with two and three cross levels (the radices of 2^17, 2^18, 2^20 and 2^21 rows; m = 8 stands in for 4096: the
    structure does not depend on m): the recursion of cross stages equals the coset DFT, and the block -> offset map is the one the
    kernels compute from the packed radices (low nibble = R_1)."""
    rnd = random.Random(11)
    m = 8
    for radices in ((2, 16), (4, 16), (16, 16), (2, 16, 16)):
        n = m
        for R in radices:
            n *= R
        logn = n.bit_length() - 1
        wN = root_of_unity(logn + 1)                  # blowup 2
        wn, s = pow(wN, 2, P), GENERATOR * wN % P      # coset 1
        a = [rnd.randrange(P) for _ in range(n)]
        y, omap = _multilevel(a, s, wn, radices, m)
        assert y == _dft([a[j] * pow(s, j, P) % P for j in range(n)], wn), radices
        # the kernels' version of the map (k_tc_lde4096_v2: out_base, k_tc_unscatter: b from o)
        packed = 0
        for l, R in enumerate(radices):
            packed |= (R.bit_length() - 1) << (4 * l)
        Pblocks = len(omap)
        prod_log = Pblocks.bit_length() - 1
        for q in range(Pblocks):                       # out_base: digits from the last level (least significant in q) to the first
            rem, off, shift = q, 0, prod_log
            for l in reversed(range(len(radices))):
                lg = (packed >> (4 * l)) & 15
                shift -= lg
                off += (rem & ((1 << lg) - 1)) << shift
                rem >>= lg
            assert off == omap[q]
        for o in range(Pblocks):                       # k_tc_unscatter: b from o, nibble by nibble
            b, oo, rd = 0, o, packed
            while rd:
                lg = rd & 15
                b = (b << lg) | (oo & ((1 << lg) - 1)); oo >>= lg; rd >>= 4
            assert omap[b] == o


def test_sharded_tree_and_leaf_index_maps():
    """The column-sharded prover's index maps (csrc/api_stark.inl) on a toy hash: k_scatter_subtrees puts the all-gathered per-rank
    sub-trees into the heap order of MerkleTree::new, k_interleave_leaves turns the all-gathered per-rank leaves (each rank's cosets
    in its own natural order) into natural order over all cosets."""
    rnd = random.Random(5)
    M = (1 << 61) - 1

    def h(a, b):
        return (a * 0x9E3779B97F4A7C15 + b * 0xC2B2AE3D27D4EB4F + 1) % M

    def heap(leaves):                                   # nodes[1] = root, nodes[count + i] would be leaf i
        count = len(leaves)
        nodes = [0] * count
        level = leaves
        lvl = count // 2
        while lvl >= 1:
            level = [h(level[2 * i], level[2 * i + 1]) for i in range(lvl)]
            nodes[lvl:2 * lvl] = level
            lvl //= 2
        return nodes

    for R in (2, 4, 8):
        count = 64 * R
        leaves = [rnd.randrange(M) for _ in range(count)]
        full = heap(leaves)
        per = count // R
        gathered = []
        for s in range(R):
            gathered += heap(leaves[s * per:(s + 1) * per])         # per entries each, entry 0 unused
        nodes = [0] * count
        for g in range(R, count):                                   # k_scatter_subtrees
            lvl = 1 << (g.bit_length() - 1)
            j, p = g - lvl, lvl // R
            s = j // p
            nodes[g] = gathered[s * per + p + (j - s * p)]
        top = heap(nodes[R:2 * R])                                  # dev_merkle(nodes + R, nodes, R)
        nodes[1:R] = top[1:R]
        assert nodes[1:] == full[1:]
    for R, B, m in ((2, 16, 8), (8, 16, 4), (4, 8, 16)):
        Bp = B // R
        natural = [rnd.randrange(M) for _ in range(B * m)]          # leaf of natural row k + B i
        gathered = []
        for s in range(R):                                          # rank s: its cosets k' < Bp in its own natural order k' + Bp i
            gathered += [natural[(s * Bp + (j % Bp)) + B * (j // Bp)] for j in range(Bp * m)]
        out = [gathered[(k // Bp) * (Bp * m) + (k % Bp) + Bp * i] for p in range(B * m) for k, i in [(p % B, p // B)]]   # k_interleave_leaves
        assert out == natural
