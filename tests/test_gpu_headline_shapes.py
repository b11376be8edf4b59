"""GPU parity at the shapes BASELINE.json's headline configs name (nothing touched them):

* C4 shape -- 68 pulsars x 10 000 TOAs, m = 72 (313 chunks of 32 TOAs, 19 level-2 flushes per work item): bins at
  both ends of the 1e6-frequency grid, interior bins and the bins ON the red-noise Fourier frequencies k/Tspan,
  per-pulsar terms of several pulsars against the longdouble truth and the oracle;
* C3 shape -- the noise-marginalised result itself (not the repo's own plain-Fp path) against the truth summed
  over all 45 pulsars;
* C5 shape -- block-diagonal N with 2500 epochs of 4 TOAs per pulsar at n = 10 000, nmfp, two draws, against the
  extended-precision truth of the equivalent GP-basis model (oracle/truth.fp_sweep_truth_blockn, pinned to the
  reference-implemented formulation in tests/test_oracle_golden.py) and the float64 oracle on the widened basis.
"""
import numpy as np
import pytest

import fastfp_b200
from conftest import EPS, term_tolerance
from fastfp_b200 import NMFP, BlockNvec, CURN_container, RN_container, synth
from oracle import fp_oracle as o
from oracle import truth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    return synth.make_config("C4")


@pytest.mark.usefixtures("sweep_path")
def test_c4_shape_terms_against_truth(c4):
    pta = c4
    assert pta.P == 68 and pta.Ts[0].shape == (10_000, 72)
    grid = synth.fp_freqs(1_000_000)
    k = np.array([1.0, 2.0, 2.5, 17.0, 30.0]) / pta.Tspan  # on / between the red-noise Fourier frequencies
    freqs = np.concatenate((grid[[0, 1, 2, 333_333, 500_000, 999_998, 999_999]], k))
    fp = fastfp_b200.FastFp(pta.psrs)
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    got = fp.per_pulsar_terms(freqs, *a)
    assert got.shape == (68, freqs.size) and np.all(np.isfinite(got))
    sel = [0, 21, 45, 67]
    args = (freqs, [pta.toas[p] for p in sel], [pta.residuals[p] for p in sel], [pta.Nvecs[p] for p in sel],
            [pta.Ts[p] for p in sel], [pta.sigmas[p] for p in sel])
    ora = o.fp_sweep(*args, per_pulsar=True)
    tt, cond = truth.fp_sweep_truth(*args)
    tol = term_tolerance(tt.astype(float), cond, ora)
    dev = np.abs(got[sel] - tt.astype(float))
    assert np.all(dev <= tol), (dev / tol).max()
    # the well-conditioned bins (above the red-noise band) meet the plain north-star tolerance
    well = freqs > 40.0 / pta.Tspan
    assert well.sum() >= 4
    assert np.abs(got[sel][:, well] / tt.astype(float)[:, well] - 1).max() < 1e-10
    # the summed statistic is the ordered pulsar sum of the terms (fastfp.py:71,90)
    acc = np.zeros(freqs.size)
    for p in range(68):
        acc = acc + got[p]
    np.testing.assert_array_equal(fp(freqs, *a), acc)


@pytest.mark.usefixtures("sweep_path")
def test_c4_shape_bins_do_not_depend_on_their_batch(c4):
    """A bin's value must not depend on which tile / launch / rank it was computed in (this is what makes the
    NCCL-gathered sweep equal to the single-GPU sweep): slices of the 1e6 grid recomputed alone, shifted by
    odd offsets so that the bins land in different tile positions."""
    pta = c4
    grid = synth.fp_freqs(1_000_000)
    fp = fastfp_b200.FastFp(pta.psrs)
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    base = fp(grid[600_000:600_000 + 9_601], *a)
    assert np.all(np.isfinite(base)) and base.min() > 0
    np.testing.assert_array_equal(fp(grid[600_000 + 37:600_000 + 37 + 1_000], *a), base[37:1_037])
    np.testing.assert_array_equal(fp(grid[600_000 + 9_600:600_000 + 9_601], *a), base[9_600:])
    # noise-only data above the red-noise band: 2 Fp ~ chi^2(2P) (examples/run_fp.ipynb:136-140)
    assert abs(2.0 * base.mean() - 2 * pta.P) < 0.15 * 2 * pta.P


@pytest.mark.usefixtures("sweep_path")
def test_c3_shape_nmfp_against_truth():
    """45 x 5000, m = 72, CURN: (draw, bin) entries of the (D, F) result against the truth of the reference
    formula with that draw's Sigma, summed over all 45 pulsars."""
    pta = synth.make_config("C3")
    curn = CURN_container(pta.Ffreqs)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    D = 11
    samples = synth.draw_samples(pta, D)
    grid = synth.nmfp_freqs(1000, pta.Tspan)
    bins = np.array([0, 1, 29, 30, 45, 500, 998, 999])
    freqs = grid[bins]
    got = NMFP(pta.psrs, sigs)(freqs, samples, pta.Nvecs, pta.Ts, pta.TNTs)
    assert got.shape == (D, bins.size)
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs, add_curn=True, curn_Ffreqs=curn.Ffreqs)
                for p, q in enumerate(pta.psrs)]
    for d in (0, 7, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, pta.TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0)), d
        well = freqs > 40.0 / pta.Tspan
        assert np.abs(got[d][well] / tv[well] - 1).max() < 1e-10
        ora = o.fp_sweep(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, sig)
        assert np.abs(got[d][well] / ora[well] - 1).max() < 1e-10


def _c5_like(P=3, n=10_000, epoch=4, seed=20240607):
    pta = synth.make_pta(P, n, seed=seed)
    rng = np.random.default_rng(seed + 777)
    blocks, tblocks, TNTs = [], [], []
    for p in range(P):
        sl = [slice(a, a + epoch) for a in range(0, n - epoch + 1, epoch)]
        jv = rng.uniform(0.3, 3.0, len(sl)) * 1e-13
        B = BlockNvec(pta.Nvecs[p], sl, jv)
        TNT = pta.Ts[p].T @ B.solve(pta.Ts[p])
        blocks.append(B)
        tblocks.append((pta.Nvecs[p], [(s.start, s.stop) for s in sl], jv))
        TNTs.append(0.5 * (TNT + TNT.T))
    return pta, blocks, tblocks, TNTs


def test_c5_shape_block_n_nmfp_two_draws():
    pta, blocks, tblocks, TNTs = _c5_like()
    assert len(blocks[0].slices) == 2500
    curn = CURN_container(pta.Ffreqs)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    D = 2
    samples = synth.draw_samples(pta, D)
    grid = synth.nmfp_freqs(10_000, pta.Tspan)
    freqs = grid[[0, 1, 30, 45, 4_999, 9_998, 9_999]]
    got = NMFP(pta.psrs, sigs)(freqs, samples, blocks, pta.Ts, TNTs)
    assert got.shape == (D, freqs.size) and np.all(np.isfinite(got))
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs, add_curn=True, curn_Ffreqs=curn.Ffreqs)
                for p, q in enumerate(pta.psrs)]
    well = freqs > 40.0 / pta.Tspan
    for d in range(D):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, TNTs, phi_args)  # TNT formed with the block N + diag(phiinv_d)
        tt, cond = truth.fp_sweep_truth_blockn(freqs, pta.toas, pta.residuals, tblocks, pta.Ts, sigmas=sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0)), d
        assert np.abs(got[d][well] / tv[well] - 1).max() < 1e-10
    # float64 oracle on the widened basis (the formulation the reference implements: epoch-indicator columns,
    # fastfp/nmfp.py:277-282), one pulsar: an LU of a 2572 x 2572 Sigma per draw
    p = 1
    n = pta.Ts[p].shape[0]
    U = np.zeros((n, 2500))
    U[np.arange(n), np.arange(n) // 4] = 1.0
    Te = np.ascontiguousarray(np.concatenate((pta.Ts[p], U), axis=1))
    pars = {k: v[0] for k, v in samples.items()}
    phi = 1.0 / o.get_phiinv(pars, **phi_args[p])
    sig_ext = Te.T @ (Te / pta.Nvecs[p][:, None]) + np.diag(1.0 / np.concatenate((phi, blocks[p].jvec)))
    one = o.fp_sweep(freqs, [pta.toas[p]], [pta.residuals[p]], [pta.Nvecs[p]], [Te], [sig_ext])
    # the engine's single-pulsar value for the same draw
    sub = {k: v[:1] for k, v in samples.items() if k.startswith(pta.psrs[p].name) or k.startswith("gw_")}
    single = NMFP([pta.psrs[p]], [sigs[p]])(freqs, sub, [blocks[p]], [pta.Ts[p]], [TNTs[p]])[0]
    assert np.abs(single[well] / one[well] - 1).max() < 1e-8  # float64 LU of the 2572-wide system: ~1e-9


def test_c5_shape_block_n_plain_fp_terms():
    pta, blocks, tblocks, _ = _c5_like(P=2, seed=99)
    sig_block = []
    for p in range(2):
        TNT = pta.Ts[p].T @ blocks[p].solve(pta.Ts[p])
        sig_block.append(0.5 * (TNT + TNT.T) + np.diag(1.0 / pta.phis[p]))
    freqs = np.concatenate((synth.fp_freqs(1_000_000)[[0, 500_000, 999_999]], np.array([1.0, 3.5]) / pta.Tspan))
    got = fastfp_b200.FastFp(pta.psrs).per_pulsar_terms(freqs, blocks, pta.Ts, sig_block)
    tt, cond = truth.fp_sweep_truth_blockn(freqs, pta.toas, pta.residuals, tblocks, pta.Ts, sigmas=sig_block)
    assert np.all(np.abs(got - tt.astype(float)) <= 1e-10 * np.abs(tt.astype(float)) + 256 * EPS * cond)
