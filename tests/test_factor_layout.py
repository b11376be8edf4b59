"""Lane-level NumPy emulation of the blocked per-draw factorisation (csrc/nmfp.cu, nmfp_factor_kernel):
the swizzled 8x8 block storage, the three fragment layouts of mma.m8n8k4.f64, the fused 8x8
Cholesky+inverse by warp shuffles, the block algorithm and the output order stage B consumes. Runs on the CPU:
it pins the index algebra the CUDA kernel is written from (the kernel itself is covered by the GPU parity tests)."""
import numpy as np
import pytest

LANE = np.arange(32)
R, Q = LANE >> 2, LANE & 3
SW = (R & 2) << 1
OFF_C = R * 8 + ((2 * Q) ^ SW)
OFF_A0 = R * 8 + (Q ^ SW)
OFF_A1 = OFF_A0 ^ 4
OFF_T0 = Q * 8 + (R ^ ((Q & 2) << 1))
OFF_T1 = OFF_T0 + 32


def dmma(d0, d1, a, b):
    """mma.m8n8k4: A[m=R][k=Q] = a[lane], B[k=Q][n=R] = b[lane], D[R][2Q], D[R][2Q+1] per lane"""
    A, B = np.zeros((8, 4)), np.zeros((4, 8))
    A[R, Q] = a
    B[Q, R] = b
    D = A @ B
    return d0 + D[R, 2 * Q], d1 + D[R, 2 * Q + 1]


def chol_inv_8x8(c0, c1):
    c0, c1 = c0.copy(), c1.copy()
    y0, y1 = (R == 2 * Q).astype(float), (R == 2 * Q + 1).astype(float)
    for j in range(8):
        sel, qs = (c1 if j & 1 else c0), j >> 1
        rinv = 1.0 / np.sqrt(sel[j * 4 + qs])
        lR = sel[(LANE & ~3) | qs] * rinv
        lc0, lc1 = sel[(8 * Q) | qs] * rinv, sel[(8 * Q + 4) | qs] * rinv
        c0 = np.where(2 * Q > j, c0 - lR * lc0, c0)
        c1 = np.where(2 * Q + 1 > j, c1 - lR * lc1, c1)
        xj0, xj1 = y0[j * 4 + Q] * rinv, y1[j * 4 + Q] * rinv
        y0 = np.where(R > j, y0 - lR * xj0, np.where(R == j, xj0, y0))
        y1 = np.where(R > j, y1 - lR * xj1, np.where(R == j, xj1, y1))
    return y0, y1


def test_access_patterns_are_bank_conflict_free():
    for off in (OFF_A0, OFF_A1, OFF_T0, OFF_T1):  # LDS.64: a half-warp covers 16 distinct 8-byte banks
        for h in (0, 1):
            assert len(set((off[16 * h:16 * h + 16] % 16).tolist())) == 16
    for qt in range(4):  # 16-byte accumulator accesses: a quarter-warp covers 8 distinct 16-byte banks
        assert len(set(((OFF_C[8 * qt:8 * qt + 8] // 2) % 8).tolist())) == 8
    assert np.all(OFF_C % 2 == 0)  # pairs stay 16-byte aligned under the swizzle


@pytest.mark.parametrize("nmbv", [1, 3, 8])
def test_blocked_factorisation_and_output_order(nmbv):
    mv = 8 * nmbv
    rng = np.random.default_rng(nmbv)
    A = rng.standard_normal((mv, mv + 5))
    S = A @ A.T + 0.1 * np.eye(mv)
    W = np.zeros(nmbv * (nmbv + 1) // 2 * 64)
    blk = lambda i, j: (i * (i + 1) // 2 + j) * 64
    ld_c = lambda b: (W[b + OFF_C].copy(), W[b + OFF_C + 1].copy())
    ld_a = lambda b: (W[b + OFF_A0].copy(), W[b + OFF_A1].copy())
    ld_t = lambda b: (W[b + OFF_T0].copy(), W[b + OFF_T1].copy())

    def st_c(b, v0, v1):
        W[b + OFF_C], W[b + OFF_C + 1] = v0, v1

    z32 = lambda: (np.zeros(32), np.zeros(32))
    for i in range(nmbv):
        for j in range(i + 1):
            st_c(blk(i, j), S[8 * i + R, 8 * j + 2 * Q], S[8 * i + R, 8 * j + 2 * Q + 1])
    for j in range(nmbv):  # left-looking block Cholesky, diagonal blocks stored inverted
        bj = [ld_a(blk(j, k)) for k in range(j)]
        c0, c1 = ld_c(blk(j, j))
        t0, t1 = z32()
        for k in range(j):
            t0, t1 = dmma(t0, t1, bj[k][0], bj[k][0])
            t0, t1 = dmma(t0, t1, bj[k][1], bj[k][1])
        st_c(blk(j, j), *chol_inv_8x8(c0 - t0, c1 - t1))
        for i in range(j + 1, nmbv):
            u0, u1 = z32()
            for k in range(j):
                a = ld_a(blk(i, k))
                u0, u1 = dmma(u0, u1, a[0], bj[k][0])
                u0, u1 = dmma(u0, u1, a[1], bj[k][1])
            v0, v1 = ld_c(blk(i, j))
            st_c(blk(i, j), v0 - u0, v1 - u1)
        bi = ld_a(blk(j, j))
        pa = [ld_a(blk(i, j)) for i in range(j + 1, nmbv)]
        for i in range(j + 1, nmbv):
            r0, r1 = dmma(*z32(), pa[i - j - 1][0], bi[0])
            st_c(blk(i, j), *dmma(r0, r1, pa[i - j - 1][1], bi[1]))
    for i in range(1, nmbv):  # X = L^-1 block row by block row, in place
        la = [ld_a(blk(i, k)) for k in range(i)]
        nx = tuple(-v for v in ld_a(blk(i, i)))
        for j in range(i):
            t0, t1 = z32()
            for k in range(j, i):
                b = ld_t(blk(k, j))
                t0, t1 = dmma(t0, t1, la[k][0], b[0])
                t0, t1 = dmma(t0, t1, la[k][1], b[1])
            st_c(blk(i, j), t0, t1)
        tb = [ld_t(blk(i, j)) for j in range(i)]
        for j in range(i):
            r0, r1 = dmma(*z32(), nx[0], tb[j][0])
            st_c(blk(i, j), *dmma(r0, r1, nx[1], tb[j][1]))
    Xr = np.linalg.inv(np.linalg.cholesky(S))
    X = np.zeros((mv, mv))
    for i in range(nmbv):
        for j in range(i + 1):
            v0, v1 = ld_c(blk(i, j))
            X[8 * i + R, 8 * j + 2 * Q], X[8 * i + R, 8 * j + 2 * Q + 1] = v0, v1
    assert np.abs(X - Xr).max() <= 64 * np.finfo(float).eps * np.abs(Xr).max() * np.linalg.cond(S) ** 0.5
    assert np.all(np.triu(X, 1) == 0.0)  # exact zeros above the diagonal (stage B multiplies full blocks)
    # output: blocks (kb, mb >= kb/2) in A-fragment order, v = X z from the same fragments
    z = rng.standard_normal(mv)
    acc = np.zeros((nmbv, 2, 32))
    nblk = 0
    for kb in range(2 * nmbv):
        for mb in range(kb // 2, nmbv):
            val = W[blk(mb, kb // 2) + (OFF_A1 if kb & 1 else OFF_A0)]
            row, col = 8 * mb + (LANE >> 2), 4 * kb + (LANE & 3)
            np.testing.assert_array_equal(val, X[row, col])
            acc[mb][0], acc[mb][1] = dmma(acc[mb][0], acc[mb][1], val, z[4 * kb + Q])
            nblk += 1
    assert nblk == nmbv * (nmbv + 1)  # linv_blocks(nmbv)
    v = np.concatenate([acc[mb][0][Q == 0] for mb in range(nmbv)])
    np.testing.assert_allclose(v, X @ z, rtol=0, atol=1e-12 * np.abs(X).max() * np.abs(z).max() * mv)
