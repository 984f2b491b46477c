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
