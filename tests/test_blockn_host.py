"""Host-side layout for a block-diagonal N (kernel ECORR): invariants of fastfp_b200.blockn (CPU)."""
import numpy as np
import pytest

from fastfp_b200 import blockn, synth


def _case(n=150, seed=0):
    pta = synth.make_pta(1, n, n_tm=4, ncomps=3, seed=seed)
    rng = np.random.default_rng(seed)
    slices, a = [], 0
    while a < n - 45:
        ln = int(rng.integers(1, 9)) if len(slices) != 3 else 41  # one epoch longer than a chunk
        slices.append(slice(a, a + ln))
        a += ln + int(rng.integers(0, 3))  # some TOAs belong to no epoch
    jvec = rng.uniform(0.2, 3.0, len(slices)) * 1e-13
    return pta, blockn.BlockNvec(pta.Nvecs[0], slices, jvec)


def test_sherman_morrison_matches_dense_inverse():
    pta, B = _case()
    x = np.random.default_rng(1).standard_normal((150, 3))
    want = np.linalg.solve(B.dense(), x)
    np.testing.assert_allclose(B.solve(x), want, rtol=1e-11)
    np.testing.assert_allclose(B.solve(x[:, 0]), want[:, 0], rtol=1e-11)


@pytest.mark.parametrize("CI", [16, 32])
def test_layout_invariants(CI):
    pta, B = _case()
    q, T = pta.psrs[0], pta.Ts[0]
    d = blockn.prepare(q.toas, q.residuals, B, T, CI)
    n2 = d["toas"].shape[0]
    assert n2 % CI == 0 and d["done_mask"].shape[0] == n2 // CI
    real = np.isfinite(d["Nvec"])
    assert real.sum() == 150  # every TOA appears exactly once, padding has infinite variance
    np.testing.assert_array_equal(np.sort(d["toas"][real]), np.sort(q.toas))
    assert np.all(d["T"][~real] == 0) and np.all(d["slot_val"][~real] == 0) and np.all(d["slot_idx"][~real] == -1)
    # every group of 4 TOAs carries at most one slot; slots are 0..7
    g = d["slot_idx"].reshape(-1, 4)
    for row in g:
        s = set(row[row >= 0])
        assert len(s) <= 1 and all(0 <= v < 8 for v in s)
    # the quadratic form x^T N^-1 y is reproduced by the diagonal part minus the folded slot sums
    rng = np.random.default_rng(2)
    x = rng.standard_normal(150)
    xs = np.zeros(n2)
    order = np.argsort(np.argsort(q.toas))  # toas are sorted and unique: map by value
    xs[real] = x[np.searchsorted(q.toas, d["toas"][real])]
    ninv = np.where(real, 1.0 / d["Nvec"], 0.0)
    diag = (xs * xs * ninv).sum()
    corr, run = 0.0, np.zeros(8)
    for c in range(n2 // CI):
        sl = slice(c * CI, (c + 1) * CI)
        for s in range(8):
            sel = d["slot_idx"][sl] == s
            run[s] += (d["slot_val"][sl][sel] * xs[sl][sel]).sum()
            if (d["done_mask"][c] >> s) & 1:
                corr += run[s] ** 2
                run[s] = 0.0
    assert np.all(run == 0.0)  # every epoch was closed
    want = x @ B.solve(x)
    assert abs((diag - corr) - want) < 1e-11 * abs(want)
    # Sherman-Morrison applied to T and r in the "(N^-1 x) * nvec" form
    Tn = B.solve(T)
    np.testing.assert_allclose((d["T"] * ninv[:, None])[real], Tn[np.searchsorted(q.toas, d["toas"][real])], rtol=1e-10, atol=1e-3)
    rn = B.solve(q.residuals)
    np.testing.assert_allclose((d["res_w"] * ninv)[real], rn[np.searchsorted(q.toas, d["toas"][real])], rtol=1e-9, atol=1e-2)


def test_plain_vector_is_a_block_n_without_epochs():
    pta = synth.make_pta(1, 70, n_tm=3, ncomps=2)
    d = blockn.prepare(pta.psrs[0].toas, pta.psrs[0].residuals, pta.Nvecs[0], pta.Ts[0], 16)
    assert np.all(d["slot_idx"] == -1) and np.all(d["done_mask"] == 0) and d["toas"].shape[0] == 80
    assert not blockn.is_block(pta.Nvecs[0]) and blockn.is_block(blockn.BlockNvec(pta.Nvecs[0], [], np.zeros(0)))


def test_bad_slices_are_rejected():
    pta = synth.make_pta(1, 40, n_tm=3, ncomps=2)
    q = pta.psrs[0]
    for sl in ([slice(0, 5), slice(3, 8)], [slice(10, 10)], [slice(30, 50)]):
        with pytest.raises(ValueError):
            blockn.prepare(q.toas, q.residuals, blockn.BlockNvec(pta.Nvecs[0], sl, np.ones(len(sl))), pta.Ts[0], 16)
