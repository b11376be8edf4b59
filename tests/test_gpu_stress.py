"""Randomised small shapes for both paths against the longdouble truth: single pulsar, a handful of
TOAs, per-draw blocks narrower than one MMA tile, single frequency / single draw, ragged everything."""
import numpy as np
import pytest

import fastfp_b200
from conftest import EPS, term_tolerance
from fastfp_b200 import NMFP, CURN_container, RN_container, synth
from oracle import fp_oracle as o
from oracle import truth

pytestmark = pytest.mark.gpu


def _shape(rng):
    P = int(rng.integers(1, 5))
    n_tm = [int(rng.integers(1, 7)) for _ in range(P)]
    ncomps = int(rng.integers(1, 13))
    # enough TOAs for the basis to be well posed, otherwise arbitrary (primes, < one chunk, ...)
    ns = [int(rng.integers(3 * (n_tm[p] + 2 * ncomps), 400)) for p in range(P)]
    return P, n_tm, ncomps, ns


@pytest.mark.usefixtures("sweep_path")
@pytest.mark.parametrize("seed", range(12))
def test_random_small_fp(seed):
    rng = np.random.default_rng(1000 + seed)
    P, n_tm, ncomps, ns = _shape(rng)
    pta = synth.make_pta(P, ns, n_tm=n_tm, ncomps=ncomps, seed=seed)
    F = int(rng.choice([1, 2, 7, 33, 65, 130]))
    freqs = np.sort(rng.uniform(2e-9, 3e-7, F))
    args = (freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, pta.sigmas)
    got = fastfp_b200.FastFp(pta.psrs).per_pulsar_terms(freqs, pta.Nvecs, pta.Ts, pta.sigmas)
    tt, cond = truth.fp_sweep_truth(*args)
    tol = term_tolerance(tt.astype(float), cond, o.fp_sweep(*args, per_pulsar=True))
    assert got.shape == (P, F)
    assert np.all(np.abs(got - tt.astype(float)) <= tol), (P, n_tm, ncomps, ns, F)


@pytest.mark.usefixtures("sweep_path")
@pytest.mark.parametrize("seed", range(12))
def test_random_small_nmfp(seed):
    rng = np.random.default_rng(2000 + seed)
    P, n_tm, ncomps, ns = _shape(rng)
    pta = synth.make_pta(P, ns, n_tm=n_tm, ncomps=ncomps, seed=seed)
    F, D = int(rng.choice([1, 3, 31, 64, 70])), int(rng.choice([1, 2, 8, 9, 17]))
    freqs = np.sort(rng.uniform(2e-9, 3e-7, F))
    samples = synth.draw_samples(pta, D, seed=seed)
    ncurn = int(rng.integers(1, ncomps + 1))
    curn = CURN_container(np.repeat(np.arange(1, ncurn + 1) / pta.Tspan, 2))
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    got = NMFP(pta.psrs, sigs)(freqs, samples, pta.Nvecs, pta.Ts, pta.TNTs)
    assert got.shape == (D, F)
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs, add_curn=True, curn_Ffreqs=curn.Ffreqs)
                for p, q in enumerate(pta.psrs)]
    for d in sorted({0, D // 2, D - 1}):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, pta.TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0)), (P, n_tm, ncomps, ns, F, D, d)
