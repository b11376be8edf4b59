"""Host digit planes of the round-2 INT8 contraction (fastfp_b200/i8planes.py): layout, exactness, offsets."""
import numpy as np
import pytest

from fastfp_b200 import i8planes as ip


@pytest.mark.parametrize("KB", [128, 64, 32])
def test_swizzle_is_a_permutation_of_each_8_row_group(KB):
    rr, cc = np.meshgrid(np.arange(ip.ROWS), np.arange(KB), indexing="ij")
    off = ip.swz_offset(rr, cc, KB)
    assert sorted(off.ravel().tolist()) == list(range(ip.ROWS * KB))
    assert np.all(off // (8 * KB) == rr // 8)            # a row stays inside its 8-row group
    assert np.all(off % 16 == cc % 16)                   # 16-byte chunks move whole
    # the formula of tools/probes/umma_i8_split_check.cu, element by element
    sh = {128: 0, 64: 1, 32: 2}[KB]
    for r, c in [(0, 0), (5, 17), (13, KB - 1), (127, 16 % KB)]:
        want = (r >> 3) * (8 * KB) + (r & 7) * KB + (((c >> 4) ^ ((r & 7) >> sh)) << 4) + (c & 15)
        assert int(ip.swz_offset(r, c, KB)) == want


def test_digits_reconstruct_to_the_fixed_point_grid():
    rng = np.random.default_rng(0)
    x = np.concatenate((rng.uniform(-1, 1, 5000), [1.0, -1.0, 0.0, 2.0 ** -40, 1 - 2.0 ** -53]))
    du = ip.unsigned_digits(x).astype(np.longdouble)
    rec = sum(du[i] * np.longdouble(2.0) ** (-7 * (i + 1)) for i in range(8))
    assert np.abs(rec - (x.astype(np.longdouble) * 0.5 + 0.5)).max() <= np.longdouble(2.0) ** -57
    assert du.max() <= 128 and du.min() >= 0 and du[1:].max() <= 127
    g = rng.standard_normal(4000) * 10.0 ** rng.uniform(-3, 3, 4000)
    e = np.ceil(np.log2(np.abs(g))) + 1
    ds = ip.signed_digits(g, e.astype(np.int64)).astype(np.longdouble)
    rec = sum(ds[i] * np.longdouble(2.0) ** (-7 * (i + 1)) for i in range(8))
    assert np.abs(rec - np.ldexp(g.astype(np.longdouble), -e.astype(np.int64))).max() <= np.longdouble(2.0) ** -57
    assert np.abs(ds).max() <= 64


@pytest.mark.parametrize("KB", [32, 128])
def test_planes_give_the_product_and_the_offset(KB):
    rng = np.random.default_rng(KB)
    m, n, F = 72, 300, 5                                   # n not a multiple of the stage
    G = rng.standard_normal((m, n)) * 10.0 ** rng.uniform(-2, 2, (m, 1))
    S = rng.uniform(-1, 1, (n, 2 * F))
    planes, e, scale, roff = ip.g_digit_planes(G, KB)
    nstage = planes.shape[0]
    assert planes.shape == (nstage, 8, ip.ROWS * KB) and nstage == -(-n // KB)
    # un-swizzle and check the digits are those of G / 2^e
    rr, cc = np.meshgrid(np.arange(ip.ROWS), np.arange(KB), indexing="ij")
    off = ip.swz_offset(rr, cc, KB)
    dG = planes[:, :, off].transpose(1, 2, 0, 3).reshape(8, ip.ROWS, nstage * KB)[:, :, :n].astype(np.int64)
    assert np.all(dG[:, m:, :] == 0)
    np.testing.assert_array_equal(dG[:, :m], ip.signed_digits(G, e[:m, None]).astype(np.int64))
    # the integer products per digit weight, recombined as the kernel would
    dS = ip.unsigned_digits(S).astype(np.int64)
    acc = [np.zeros((ip.ROWS, 2 * F), dtype=np.int64) for _ in range(8)]
    for i in range(8):
        for j in range(8 - i):
            acc[i + j] += dG[i] @ dS[j]
    assert max(np.abs(a).max() for a in acc) < 2 ** 31
    y = np.zeros((ip.ROWS, 2 * F))
    for g in range(7, -1, -1):
        y += acc[g].astype(np.float64) * 2.0 ** (-7 * (g + 2))
    Y = (y - roff[:, None]) * scale[:, None]
    Yt = (G.astype(np.longdouble) @ S.astype(np.longdouble)).astype(np.float64)
    bound = np.finfo(float).eps * (np.abs(G) @ np.abs(S))
    assert np.all(np.abs(Y[:m] - Yt) <= bound) and np.all(Y[m:] == 0)
    # roff is (1/2) sum_i G_ji / 2^e
    want = 0.5 * np.ldexp(G.astype(np.longdouble), -e[:m, None]).sum(axis=1)
    assert np.abs(roff[:m] - want.astype(np.float64)).max() <= 4 * np.finfo(float).eps * np.abs(want).max() + 2.0 ** -50
