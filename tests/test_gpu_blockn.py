"""GPU parity of the block-diagonal-N (kernel ECORR) path. The reference has no implementation of it
(fastfp/utils.py:29-31); parity is pinned through the mathematically identical GP-basis formulation the
reference does implement: C = D + U J U^T + T Phi T^T with U the epoch-indicator matrix as extra basis
columns, evaluated by the oracle and the longdouble truth."""
import numpy as np
import pytest

import fastfp_b200
from conftest import EPS, Psr
from fastfp_b200 import NMFP, BlockNvec, RN_container, synth
from oracle import fp_oracle as o
from oracle import truth

pytestmark = pytest.mark.gpu


def _epochs(n, rng, sizes=(1, 9), gap=(0, 2), long_at=None):
    slices, a = [], 0
    while a < n - 70:
        ln = 70 if long_at is not None and len(slices) == long_at else int(rng.integers(*sizes))
        slices.append(slice(a, a + ln))
        a += ln + int(rng.integers(*gap))
    return slices


def _build(P=3, ns=(400, 613, 300), n_tm=(8, 12, 10), ncomps=30, seed=17):
    pta = synth.make_pta(P, list(ns), n_tm=list(n_tm), ncomps=ncomps, seed=seed)
    rng = np.random.default_rng(seed)
    blocks, Text, phiext, sig_block, TNT_block = [], [], [], [], []
    for p in range(P):
        n = ns[p]
        sl = _epochs(n, rng, long_at=2 if p == 1 else None) if p != 2 else []  # pulsar 2: plain diagonal N
        jv = rng.uniform(0.3, 3.0, len(sl)) * 1e-13
        B = BlockNvec(pta.Nvecs[p], sl, jv)
        blocks.append(B if p != 2 else pta.Nvecs[p])
        T, phi = pta.Ts[p], pta.phis[p]
        U = np.zeros((n, len(sl)))
        for e, s in enumerate(sl):
            U[s, e] = 1.0
        Text.append(np.ascontiguousarray(np.concatenate((T, U), axis=1)))
        phiext.append(np.concatenate((phi, jv)))
        TNT = T.T @ B.solve(T)
        TNT = 0.5 * (TNT + TNT.T)
        TNT_block.append(TNT)
        sig_block.append(TNT + np.diag(1.0 / phi))
    sig_ext = [Te.T @ (Te / pta.Nvecs[p][:, None]) + np.diag(1.0 / phiext[p]) for p, Te in enumerate(Text)]
    return pta, blocks, Text, sig_ext, sig_block, TNT_block


def test_block_n_fp_matches_gp_basis_formulation():
    pta, blocks, Text, sig_ext, sig_block, _ = _build()
    freqs = np.concatenate((synth.fp_freqs(60), np.array([1.0, 3.5]) / pta.Tspan))
    got = fastfp_b200.FastFp(pta.psrs)(freqs, blocks, pta.Ts, sig_block)
    args = (freqs, pta.toas, pta.residuals, pta.Nvecs, Text, sig_ext)
    want = o.fp_sweep(*args)
    tt, cond = truth.fp_sweep_truth(*args)
    tv = tt.sum(0).astype(float)
    assert got.shape == want.shape
    assert np.all(np.abs(got - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0))
    assert np.abs(got / want - 1).max() < 1e-6
    # and it is not the diagonal-N answer: ECORR matters
    plain = fastfp_b200.FastFp(pta.psrs)(freqs, pta.Nvecs, pta.Ts, pta.sigmas)
    assert np.abs(plain / tv - 1).max() > 1e-3


def test_block_n_nmfp_matches_gp_basis_formulation():
    pta, blocks, Text, _, _, TNT_block = _build(seed=23)
    D, F = 6, 37
    samples = synth.draw_samples(pta, D)
    freqs = synth.nmfp_freqs(F, pta.Tspan) * 1.0071
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs) for q in pta.psrs]
    got = NMFP(pta.psrs, sigs)(freqs, samples, blocks, pta.Ts, TNT_block)
    assert got.shape == (D, F)
    for d in (0, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig_ext = []
        for p, q in enumerate(pta.psrs):
            phi = sigs[p].update_phi(pars)
            ne = Text[p].shape[1] - pta.Ts[p].shape[1]
            jv = np.asarray(blocks[p]._jvec) if ne else np.zeros(0)
            phie = np.concatenate((phi, jv))
            Te = Text[p]
            sig_ext.append(Te.T @ (Te / pta.Nvecs[p][:, None]) + np.diag(1.0 / phie))
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, Text, sig_ext)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0)), d


def test_get_xcy_with_block_n():
    rng = np.random.default_rng(4)
    n, m = 180, 7
    T, nvec = rng.standard_normal((n, m)), rng.uniform(0.5, 2.0, n)
    sl = [slice(0, 5), slice(5, 7), slice(40, 90)]
    B = BlockNvec(nvec, sl, np.array([0.7, 1.3, 0.2]))
    phi = rng.uniform(0.1, 3.0, m)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    sigma = T.T @ B.solve(T) + np.diag(1 / phi)
    C = B.dense() + T @ np.diag(phi) @ T.T  # dense known answer
    want = x @ np.linalg.solve(C, y)
    got = fastfp_b200.get_xCy(B, T, sigma, x, y)
    assert abs(got - want) < 1e-11 * abs(want) + 1e-12
