"""The oracle against the golden vectors produced by the reference's own source files
(tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

from conftest import EPS, term_tolerance
from oracle import fp_oracle as o
from oracle import truth


@pytest.mark.parametrize("name", ["fp_white", "fp_red"])
def test_literal_oracle_reproduces_reference_source(golden, name):
    g = golden(name)
    lit = np.array(
        [o.calculate_Fp(f, g.lst("toas"), g.lst("res"), g.lst("Nvec"), g.lst("T"), g.lst("sigma")) for f in g["freqs"]]
    )
    # same formulas, same operation order, same NumPy primitives -> identical to the last bit
    np.testing.assert_array_equal(lit, g["ref_fp"])


@pytest.mark.parametrize("name", ["fp_white", "fp_red"])
def test_batched_oracle_within_conditioning_of_truth(golden, name):
    g = golden(name)
    args = (g["freqs"], g.lst("toas"), g.lst("res"), g.lst("Nvec"), g.lst("T"), g.lst("sigma"))
    bat = o.fp_sweep(*args, per_pulsar=True)
    tol = term_tolerance(g["truth_terms"], g["cond"], bat, k_oracle=1.0)  # E_p calibrates itself
    assert np.all(np.abs(bat - g["truth_terms"]) <= tol)
    # and the summed sweep agrees with the reference's own sum to the summed allowance
    allow = 1e-10 * np.abs(g["ref_fp"]) + 64 * EPS * g["cond"].sum(0)
    assert np.all(np.abs(o.fp_sweep(*args) - g["ref_fp"]) <= allow)


def test_get_xcy_golden(golden):
    g = golden("fp_white")
    for p in range(g.P):
        v = o.get_xCy(g[f"Nvec_{p}"], g[f"T_{p}"], g[f"sigma_{p}"], g[f"x_{p}"], g[f"y_{p}"])
        assert v == g["ref_xcy"][p]


def test_truth_matches_dense_covariance():
    """Known answer independent of Woodbury: build C = N + T diag(phi) T^T densely (small n,
    moderate phi so C is well conditioned) and compare x^T C^-1 y."""
    rng = np.random.default_rng(3)
    n, m = 60, 7
    T = rng.standard_normal((n, m))
    Nvec = rng.uniform(0.5, 2.0, n)
    phi = rng.uniform(0.1, 3.0, m)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    C = np.diag(Nvec) + T @ np.diag(phi) @ T.T
    dense = x @ np.linalg.solve(C, y)
    sigma = T.T @ (T / Nvec[:, None]) + np.diag(1 / phi)
    assert abs(o.get_xCy(Nvec, T, sigma, x, y) - dense) < 1e-12 * abs(dense) + 1e-13
    tv, _ = truth.get_xCy_truth(Nvec, T, sigma, x, y)
    assert abs(float(tv) - dense) < 1e-12 * abs(dense) + 1e-13


def test_exact_signal_identity(golden):
    """If r = a*s + b*c then N = M theta and Fp = 0.5 theta^T M theta (SURVEY 8c KAT)."""
    g = golden("fp_red")
    p, f = 1, g["freqs"][9]
    toa, Nvec, T, sigma = g["toas_1"], g["Nvec_1"], g["T_1"], g["sigma_1"]
    s, c = np.sin(2 * np.pi * f * toa), np.cos(2 * np.pi * f * toa)
    r = 0.3 * s - 1.1 * c
    fp = o.calculate_Fp(f, [toa], [r], [Nvec], [T], [sigma])
    pref = 1 / f ** (1 / 3)
    A = pref * np.stack((s, c))
    M = np.array([[o.get_xCy(Nvec, T, sigma, A[i], A[j]) for j in range(2)] for i in range(2)])
    th = np.array([0.3, -1.1]) / pref
    assert abs(fp - 0.5 * th @ M @ th) < 1e-8 * abs(fp)


def test_powerlaw_and_phi_layouts(golden):
    g = golden("nmfp")
    pars = {k[len("sample__"):]: g[k][0] for k in g.g.files if k.startswith("sample__")}
    Ff, Fc = g["Ffreqs"], g["Ffreqs_curn"]
    ntm = [int(g[f"ntm_{p}"]) for p in range(g.P)]
    names = [str(g[f"name_{p}"]) for p in range(g.P)]
    np.testing.assert_array_equal(o.powerlaw(Fc, pars["gw_log10_A"], pars["gw_gamma"]), g["ref_curn_phi0"])
    plain = np.concatenate([o.get_phi(pars, names[p], ntm[p], Ff) for p in range(g.P)])
    np.testing.assert_array_equal(plain, g["ref_phi0_plain"])
    curn = np.concatenate(
        [o.get_phiinv(pars, names[p], ntm[p], Ff, add_curn=True, curn_Ffreqs=Fc) for p in range(g.P)]
    )
    np.testing.assert_array_equal(curn, g["ref_phiinv0"])
    ec = o.ecorr_phi([np.ones(int(k)) for k in g["ecorr_nw"]], g["ecorr_log10"])
    np.testing.assert_array_equal(ec, g["ref_ecorr_phi"])
    np.testing.assert_array_equal(o.get_phi(pars, names[0], ntm[0], Ff, ecorr_phi_fixed=ec), g["ref_phi_tm_ecorr_rn"])
    np.testing.assert_array_equal(
        o.get_phi(pars, names[0], ntm[0], Ff, add_curn=True, curn_Ffreqs=Fc, ecorr_phi_fixed=ec),
        g["ref_phi_tm_ecorr_rn_curn"],
    )


def _phi_args(g, curn):
    out = []
    for p in range(g.P):
        kw = dict(psr_name=str(g[f"name_{p}"]), n_tm=int(g[f"ntm_{p}"]), Ffreqs=g["Ffreqs"])
        if curn:
            kw.update(add_curn=True, curn_Ffreqs=g["Ffreqs_curn"])
        out.append(kw)
    return out


def test_get_sigmas_and_nmfp_goldens(golden):
    g = golden("nmfp")
    samples = {k[len("sample__"):]: g[k] for k in g.g.files if k.startswith("sample__")}
    pars0 = {k: v[0] for k, v in samples.items()}
    for p, s in enumerate(o.get_sigmas(pars0, g.lst("TNT"), _phi_args(g, True))):
        np.testing.assert_array_equal(s, g[f"ref_sigma0_{p}"])
    D = int(g["D"])
    for curn, key in ((True, "ref_nmfp_curn"), (False, "ref_nmfp_plain")):
        pa = _phi_args(g, curn)
        lit = np.array(
            [
                [
                    o.calculate_nmfp(f, {k: v[d] for k, v in samples.items()}, g.lst("toas"), g.lst("res"),
                                     g.lst("Nvec"), g.lst("T"), g.lst("TNT"), pa)
                    for f in g["freqs"]
                ]
                for d in range(D)
            ]
        )
        np.testing.assert_array_equal(lit, g[key])  # (D, F), draw-major
    bat = o.nmfp_sweep(g["freqs"], samples, g.lst("toas"), g.lst("res"), g.lst("Nvec"), g.lst("T"), g.lst("TNT"),
                       _phi_args(g, True))
    allow = 1e-10 * np.abs(g["truth_nmfp_curn"]) + 256 * EPS * g["cond_curn"]
    assert np.all(np.abs(bat - g["truth_nmfp_curn"]) <= allow)


def test_block_n_truth_equals_the_gp_basis_truth():
    """oracle/truth.fp_sweep_truth_blockn (Sherman-Morrison in longdouble) is the same quantity as the
    reference-implemented GP-basis model (epoch-indicator columns appended to T, fastfp/nmfp.py:277-282)
    evaluated by fp_sweep_truth: pinned here at a size where the O((m + n_epoch)^3) form runs."""
    from fastfp_b200 import synth
    from oracle import truth

    pta = synth.make_pta(2, [203, 160], n_tm=[4, 5], ncomps=6, seed=3)
    rng = np.random.default_rng(1)
    blocks, Text, sig_ext, phiinvs = [], [], [], []
    for p in range(2):
        n = pta.psrs[p].toas.size
        sl = [(a, a + 4) for a in range(0, n - 40, 4)] + [(n - 30, n - 21)]
        jv = rng.uniform(0.3, 3.0, len(sl)) * 1e-13
        blocks.append((pta.Nvecs[p], sl, jv))
        U = np.zeros((n, len(sl)))
        for e, (a, b) in enumerate(sl):
            U[a:b, e] = 1.0
        Te = np.concatenate((pta.Ts[p], U), axis=1)
        Text.append(Te)
        LDt = np.longdouble
        sig = (Te.astype(LDt).T @ (Te.astype(LDt) / pta.Nvecs[p].astype(LDt)[:, None])
               + np.diag(1 / np.concatenate((pta.phis[p], jv)).astype(LDt)))
        sig_ext.append(sig)
        phiinvs.append(1.0 / pta.phis[p])
    freqs = np.concatenate((synth.fp_freqs(7), np.array([1.0, 2.5]) / pta.Tspan))
    tb, cb = truth.fp_sweep_truth_blockn(freqs, pta.toas, pta.residuals, blocks, pta.Ts, phiinvs)
    tg, cg = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, Text, sig_ext)
    # two longdouble evaluations of one quantity: they agree to a small multiple of the LONGDOUBLE rounding
    # times the conditioning figure (which is in units of the relative rounding error)
    eps_ld = float(np.finfo(np.longdouble).eps)
    assert np.all(np.abs((tb - tg).astype(float)) <= 1e-14 * np.abs(tg.astype(float)) + 1e4 * eps_ld * np.maximum(cb, cg))
