"""Fe-statistic (the reference's to-do, README.md:23) on the GPU against oracle/fe_oracle.py -- the published statistic
restated on top of the reference's own get_xCy -- for several sky positions, on both sweep kernels."""
import numpy as np
import pytest

import fastfp_b200
from fastfp_b200 import synth
from fastfp_b200.fe import antenna_pattern
from oracle import fe_oracle

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("sweep_path")]


def test_fe_matches_the_oracle_over_a_sky_grid():
    pta = synth.make_pta(4, [300, 257, 411, 350], n_tm=[6, 8, 5, 7], ncomps=10, seed=12)
    freqs = np.concatenate((synth.fp_freqs(40)[8::4], np.array([2.5, 7.0]) / pta.Tspan))
    th = np.array([0.3, 1.2, 2.6])
    ph = np.array([0.1, 3.0, 5.5])
    fe = fastfp_b200.FastFe(pta.psrs)
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    got = fe.calculate_Fe(freqs, th, ph, *a)
    assert got.shape == (3, freqs.size) and np.all(np.isfinite(got))
    pos = [q.pos for q in pta.psrs]
    want = np.array([[fe_oracle.calculate_Fe(f, t, p_, pta.toas, pta.residuals, pos, *a) for f in freqs]
                     for t, p_ in zip(th, ph)])
    well = freqs > 40.0 / pta.Tspan
    assert well.sum() >= 6
    assert np.abs(got[:, well] / want[:, well] - 1).max() < 1e-9
    assert np.abs(got / want - 1).max() < 1e-5  # bins inside the red-noise band: ill-conditioned in the formula itself
    # batching forms
    # (NumPy's vectorised sin/cos of the sky-angle ARRAYS may differ from the scalar calls by an ulp, so the forms that
    # change how the antenna patterns are evaluated are compared to 1e-13; the frequency batching is bit for bit: a bin
    # does not depend on how many bins are swept with it)
    one = fe.calculate_Fe(float(freqs[3]), float(th[1]), float(ph[1]), *a)
    assert np.ndim(one) == 0 and abs(one / got[1, 3] - 1) < 1e-13
    np.testing.assert_allclose(fe.calculate_Fe(freqs, float(th[2]), float(ph[2]), *a), got[2], rtol=1e-13)
    np.testing.assert_array_equal(fe.calculate_Fe(float(freqs[0]), th, ph, *a), got[:, 0])
    # the same pack serves Fp: a single pulsar seen with F+ = 1, Fx = 0 would reduce Fe to its Fp term; here just
    # check that the Fp sweep of the shared pack is unaffected by the Fe call
    fp = fastfp_b200.FastFp(pta.psrs)(freqs, *a)
    np.testing.assert_array_equal(fe.calculate_Fp(freqs, *a), fp)


def test_fe_device_resident_and_quadratic_scaling():
    import torch
    from conftest import Psr

    pta = synth.make_pta(3, 500, n_tm=8, ncomps=12, seed=3)
    f = synth.fp_freqs(70)
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    fe = fastfp_b200.FastFe(pta.psrs)
    th, ph = np.array([0.7, 2.0]), np.array([1.0, 4.0])
    host = fe.calculate_Fe(f, th, ph, *a)
    dev = fe.calculate_Fe(torch.tensor(f, dtype=torch.float64, device="cuda"), th, ph, *a)
    np.testing.assert_array_equal(dev.cpu().numpy(), host)
    psr2 = [Psr(q.toas, 2.0 * q.residuals) for q in pta.psrs]
    for q2, q in zip(psr2, pta.psrs):
        q2.pos = q.pos
    np.testing.assert_array_equal(fastfp_b200.FastFe(psr2).calculate_Fe(f, th, ph, *a), 4.0 * host)


def test_antenna_pattern_matches_the_oracle_definition():
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((5, 3))
    pos /= np.linalg.norm(pos, axis=1, keepdims=True)
    th, ph = rng.uniform(0, np.pi, 7), rng.uniform(0, 2 * np.pi, 7)
    fp, fx = antenna_pattern(pos, th, ph)
    assert fp.shape == fx.shape == (7, 5)
    for k in range(7):
        for p in range(5):
            w = fe_oracle.antenna_pattern(pos[p], th[k], ph[k])
            assert abs(fp[k, p] - w[0]) < 1e-15 and abs(fx[k, p] - w[1]) < 1e-15
