"""GPU parity tests for the noise-marginalised path (NMFP.calculate_nmfp through the C ABI)."""
import numpy as np
import pytest

import fastfp_b200
from conftest import EPS, Psr
from fastfp_b200 import NMFP, CURN_container, GPEcorr_container, RN_container, synth
from oracle import fp_oracle as o
from oracle import truth

pytestmark = pytest.mark.gpu


def _samples(g):
    return {k[len("sample__"):]: g[k] for k in g.g.files if k.startswith("sample__")}


def _tol(truth_vals, cond):
    # the nmfp grid sits exactly on the red-noise Fourier frequencies (run_nmfp.py:247): the
    # worst-conditioned points there are. 1e-10 relative plus the conditioning allowance.
    return 1e-10 * np.abs(truth_vals) + 256 * EPS * cond


@pytest.mark.usefixtures("sweep_path")
def test_nmfp_matches_reference_goldens(golden):
    g = golden("nmfp")
    samples = _samples(g)
    mats = (g.lst("Nvec"), g.lst("T"), g.lst("TNT"))
    curn = CURN_container(g["Ffreqs_curn"])
    sig_c = [RN_container(q, Ffreqs=g["Ffreqs"], add_curn=True, curn_container=curn) for q in g.psrs]
    nm = NMFP(g.psrs, sig_c)
    got = nm(g["freqs"], samples, *mats)
    assert got.shape == g["ref_nmfp_curn"].shape  # (D, F), draw-major
    tol = _tol(g["truth_nmfp_curn"], g["cond_curn"])
    assert np.all(np.abs(got - g["truth_nmfp_curn"]) <= tol)
    assert np.all(np.abs(got - g["ref_nmfp_curn"]) <= 2 * tol)
    # the nested-vmap spelling of examples/run_nmfp.py:265-266
    vf = fastfp_b200.vmap(nm, in_axes=(0, None, None, None, None))
    vg = fastfp_b200.vmap(vf, in_axes=(None, 0, None, None, None))
    np.testing.assert_array_equal(vg(g["freqs"], samples, *mats), got)
    # without the common process
    nm2 = NMFP(g.psrs, [RN_container(q, Ffreqs=g["Ffreqs"]) for q in g.psrs])
    got2 = nm2(g["freqs"], samples, *mats)
    assert np.all(np.abs(got2 - g["ref_nmfp_plain"]) <= 2 * tol + 1e-9 * np.abs(g["ref_nmfp_plain"]))


def test_nmfp_batching_shapes(golden):
    g = golden("nmfp")
    samples = _samples(g)
    mats = (g.lst("Nvec"), g.lst("T"), g.lst("TNT"))
    nm = NMFP(g.psrs, [RN_container(q, Ffreqs=g["Ffreqs"]) for q in g.psrs])
    full = nm(g["freqs"], samples, *mats)
    pars1 = {k: v[1] for k, v in samples.items()}
    row = nm(g["freqs"], pars1, *mats)
    assert row.shape == (g["freqs"].shape[0],)
    np.testing.assert_array_equal(row, full[1])
    colv = nm(float(g["freqs"][2]), samples, *mats)
    assert colv.shape == (int(g["D"]),)
    np.testing.assert_array_equal(colv, full[:, 2])
    one = nm(float(g["freqs"][2]), pars1, *mats)
    assert np.ndim(one) == 0 and one == full[1, 2]


def test_device_powerlaw_matches_host_containers(golden):
    import torch

    g = golden("nmfp")
    samples = _samples(g)
    curn = CURN_container(g["Ffreqs_curn"])
    sigs = [RN_container(q, Ffreqs=g["Ffreqs"], add_curn=True, curn_container=curn) for q in g.psrs]
    nm = NMFP(g.psrs, sigs)
    pack = nm.prepare(g.lst("Nvec"), g.lst("T"), g.lst("TNT"))
    D = int(g["D"])
    A = np.stack([samples[s.rn_A_name] for s in sigs], axis=1)
    G = np.stack([samples[s.rn_gam_name] for s in sigs], axis=1)
    out = torch.empty((D, pack.mvar_total), dtype=torch.float64, device="cuda")
    pack.powerlaw_phiinv([s.Ffreqs for s in sigs], A, G, curn.Ffreqs, samples["gw_log10_A"], samples["gw_gamma"],
                         out.data_ptr())
    torch.cuda.synchronize()
    host = np.concatenate([s.get_phiinv(samples)[:, s.tm_weights.shape[0]:] for s in sigs], axis=1)
    np.testing.assert_allclose(out.cpu().numpy(), host, rtol=4e-15)  # pow() differs from NumPy's by ulps


def _ecorr_case(seed=11):
    """A pulsar set whose basis carries GP-ECORR columns: T = [tm | epoch indicators | Fourier]."""
    pta = synth.make_pta(3, [240, 320, 200], n_tm=[8, 10, 6], ncomps=30, seed=seed)
    psrs, Ts, TNTs, sigs, phi_args = [], [], [], [], []
    wn = {}
    for p, q in enumerate(pta.psrs):
        n, ntm = q.toas.size, pta.n_tm[p]
        nep = n // 4
        U = np.zeros((n, nep))
        U[np.arange(nep * 4), np.repeat(np.arange(nep), 4)] = 1.0
        T = np.concatenate((pta.Ts[p][:, :ntm], U, pta.Ts[p][:, ntm:]), axis=1)
        TNT = T.T @ (T / pta.Nvecs[p][:, None])
        flags = np.array(["X"] * n)
        psr = Psr(q.toas, q.residuals, name=q.name, Mmat=np.zeros((n, ntm)), backend_flags=flags)
        wn[f"{q.name}_basis_ecorr_X_log10_ecorr"] = -6.5 - 0.2 * p
        ec = GPEcorr_container(psr, [np.ones(nep)], fix_wn_vals=wn)
        sigs.append(RN_container(psr, Ffreqs=pta.Ffreqs, gp_ecorr=True, ecorr_container=ec))
        phi_args.append(dict(psr_name=q.name, n_tm=ntm, Ffreqs=pta.Ffreqs, ecorr_phi_fixed=ec.get_phi({})))
        psrs.append(psr); Ts.append(T); TNTs.append(0.5 * (TNT + TNT.T))
    return pta, psrs, Ts, TNTs, sigs, phi_args


@pytest.mark.usefixtures("sweep_path")
def test_gp_ecorr_columns_are_eliminated_as_fixed_block():
    pta, psrs, Ts, TNTs, sigs, phi_args = _ecorr_case()
    D, F = 5, 40
    samples = synth.draw_samples(pta, D)
    freqs = synth.nmfp_freqs(F, pta.Tspan) * 1.013  # off the exact Fourier grid: well conditioned
    nm = NMFP(psrs, sigs)
    got = nm(freqs, samples, pta.Nvecs, Ts, TNTs)
    toas, res = [q.toas for q in psrs], [q.residuals for q in psrs]
    want = o.nmfp_sweep(freqs, samples, toas, res, pta.Nvecs, Ts, TNTs, phi_args)
    assert got.shape == (D, F)
    # truth for one draw
    pars = {k: v[2] for k, v in samples.items()}
    tt, cond = truth.fp_sweep_truth(freqs, toas, res, pta.Nvecs, Ts, o.get_sigmas(pars, TNTs, phi_args))
    assert np.all(np.abs(got[2] - tt.sum(0).astype(float)) <= _tol(tt.sum(0).astype(float), cond.sum(0)))
    assert np.abs(got / want - 1).max() < 1e-7


@pytest.mark.usefixtures("sweep_path")
@pytest.mark.parametrize("ncomps,P", [(10, 2), (30, 5), (45, 3), (60, 2), (64, 2)])
def test_nmfp_against_oracle_shapes(ncomps, P):
    pta = synth.make_pta(P, [300 + 57 * p for p in range(P)], n_tm=[6 + p for p in range(P)], ncomps=ncomps, seed=31)
    D, F = 11, 45  # not multiples of the draw tile (8) or the frequency tile (32)
    samples = synth.draw_samples(pta, D)
    freqs = np.concatenate((synth.nmfp_freqs(5, pta.Tspan), synth.fp_freqs(F - 5)))
    curn = CURN_container(np.repeat(np.arange(1, 6) / pta.Tspan, 2))
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    got = NMFP(pta.psrs, sigs)(freqs, samples, pta.Nvecs, pta.Ts, pta.TNTs)
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs, add_curn=True, curn_Ffreqs=curn.Ffreqs)
                for p, q in enumerate(pta.psrs)]
    for d in (0, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, pta.TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= _tol(tv, cond.sum(0))), (ncomps, d)


def test_nmfp_consistent_with_plain_fp():
    """With the draw equal to the fixed noise values the two paths must agree (same Sigma)."""
    pta = synth.make_pta(3, 400, n_tm=12, ncomps=30, seed=8, inc_cp=False)
    freqs = synth.fp_freqs(50)
    fp = fastfp_b200.FastFp(pta.psrs)(freqs, pta.Nvecs, pta.Ts, pta.sigmas)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs) for q in pta.psrs]
    nm = NMFP(pta.psrs, sigs)(freqs, pta.noise, pta.Nvecs, pta.Ts, pta.TNTs)
    well = freqs > 40 / pta.Tspan
    assert np.abs(nm[well] / fp[well] - 1).max() < 1e-9
    assert np.abs(nm / fp - 1).max() < 1e-5


@pytest.mark.usefixtures("sweep_path")
def test_nmfp_full_size_properties():
    """BASELINE configs[2] shapes (45 pulsars x 5000 TOAs, m = 72): properties that do not need the oracle
    at every bin -- draw batching and draw order do not change a value, a draw equal to the fixed noise
    values reproduces the plain-Fp path (same Sigma), r -> 2r scales exactly by 4."""
    pta = synth.make_config("C3")
    curn = CURN_container(pta.Ffreqs)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    nm = NMFP(pta.psrs, sigs)
    mats = (pta.Nvecs, pta.Ts, pta.TNTs)
    D, F = 19, 150  # ragged: not multiples of the 8-draw / 64-frequency CTA tiles
    samples = synth.draw_samples(pta, D)
    freqs = synth.fp_freqs(10_000)[:: 10_000 // F][:F]
    full = nm(freqs, samples, *mats)
    assert full.shape == (D, F) and np.all(np.isfinite(full)) and full.min() > 0
    perm = np.random.default_rng(0).permutation(D)
    np.testing.assert_array_equal(nm(freqs, {k: v[perm] for k, v in samples.items()}, *mats), full[perm])
    np.testing.assert_array_equal(nm(freqs, {k: v[5:9] for k, v in samples.items()}, *mats), full[5:9])
    np.testing.assert_array_equal(nm(freqs[37:101], samples, *mats), full[:, 37:101])
    psr2 = [Psr(q.toas, 2.0 * q.residuals) for q in pta.psrs]
    np.testing.assert_array_equal(NMFP(psr2, sigs)(freqs, samples, *mats), 4.0 * full)
    # plain-Fp path with the Sigma of draw 3, built on the host the way NMFP._get_sigmas does
    pars = {k: v[3] for k, v in samples.items()}
    fp = fastfp_b200.FastFp(pta.psrs)(freqs, pta.Nvecs, pta.Ts, nm._get_sigmas(pars, pta.TNTs))
    well = freqs > 40 / pta.Tspan
    assert np.abs(full[3][well] / fp[well] - 1).max() < 1e-9
    assert np.abs(full[3] / fp - 1).max() < 1e-5


@pytest.mark.usefixtures("sweep_path")
def test_nmfp_mixed_per_draw_widths_in_one_pack():
    """Pulsars with different numbers of red-noise components share one pack: the per-draw blocks are
    padded to the widest one (each at its own top-left offset) and stage B skips a different number of
    leading k-blocks per pulsar."""
    import copy

    parts = [synth.make_pta(2, [260, 301], n_tm=[3, 5], ncomps=5, seed=41, inc_cp=False),
             synth.make_pta(2, [333, 280], n_tm=[4, 2], ncomps=14, seed=42, inc_cp=False)]
    psrs, Nvecs, Ts, TNTs, sigs, phi_args = [], [], [], [], [], []
    for k, pta in enumerate(parts):
        for p, q in enumerate(pta.psrs):
            q = copy.copy(q)
            q.name = f"{q.name}_{k}"  # the two synthetic sets reuse names
            psrs.append(q)
            Nvecs.append(pta.Nvecs[p]); Ts.append(pta.Ts[p]); TNTs.append(pta.TNTs[p])
            sigs.append(RN_container(q, Ffreqs=pta.Ffreqs))
            phi_args.append(dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs))
    D, F = 9, 37
    rng = np.random.default_rng(3)
    samples = {}
    for q in psrs:
        samples[f"{q.name}_red_noise_log10_A"] = rng.uniform(-15.0, -13.0, D)
        samples[f"{q.name}_red_noise_gamma"] = rng.uniform(1.0, 6.0, D)
    freqs = np.sort(rng.uniform(2e-9, 3e-7, F))
    got = NMFP(psrs, sigs)(freqs, samples, Nvecs, Ts, TNTs)
    assert got.shape == (D, F)
    toas, res = [q.toas for q in psrs], [q.residuals for q in psrs]
    for d in (0, 4, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, toas, res, Nvecs, Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= _tol(tv, cond.sum(0))), d


def test_nmfp_wide_timing_model_block():
    """Basis wider than 320 columns (many timing-model / DMX columns): the draw-independent
    block is eliminated once per pulsar as usual, the sweep runs on the 640-row kernel family."""
    pta = synth.make_pta(2, [2600, 3001], n_tm=[330, 400], ncomps=12, seed=19)
    assert [T.shape[1] for T in pta.Ts] == [354, 424]
    D, F = 3, 21
    samples = synth.draw_samples(pta, D)
    freqs = np.concatenate((synth.nmfp_freqs(4, pta.Tspan), synth.fp_freqs(F - 4)))
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs) for q in pta.psrs]
    got = NMFP(pta.psrs, sigs)(freqs, samples, pta.Nvecs, pta.Ts, pta.TNTs)
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=pta.Ffreqs) for p, q in enumerate(pta.psrs)]
    for d in (0, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, pta.TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= _tol(tv, cond.sum(0))), d


def test_stage_a_blocks_plus_stage_b_equal_the_combined_sweep():
    """The two halves of the sweep through the C ABI, stage A made in two frequency slices laid out as two blocks
    (what two ranks would all-gather: fastfp_b200/parallel.py::tile_blocks), stage B on the blocks: bit for bit the
    combined call. One GPU stands in for the two ranks."""
    import torch

    from fastfp_b200 import parallel

    pta = synth.make_pta(4, [700, 1203, 333, 901], n_tm=[8, 12, 5, 10], ncomps=30, seed=21)
    curn = CURN_container(pta.Ffreqs)
    sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
    nm = NMFP(pta.psrs, sigs)
    mats = (pta.Nvecs, pta.Ts, pta.TNTs)
    F, D, world = 173, 11, 2  # 6 tiles: blocks of 4 (2 real + padding in the second)
    f = torch.from_numpy(synth.nmfp_freqs(F, pta.Tspan) * 1.001).cuda()
    samples = synth.draw_samples(pta, D)
    want = nm(f, samples, *mats)
    pack = nm.prepare(*mats)
    nt, per = parallel.tile_blocks(F, world)
    assert (nt, per) == (6, 4)
    zt, at = pack.nmfp_tile_sizes()
    zall = torch.full((world * per * zt,), float("nan"), dtype=torch.float64, device="cuda")
    aall = torch.full((world * per * at,), float("nan"), dtype=torch.float64, device="cuda")
    for r in range(world):
        idx = torch.arange(32 * r * per, 32 * (r + 1) * per, device="cuda").clamp_(max=F - 1)
        floc = f[idx].contiguous()
        pack.nmfp_stage_a(floc.data_ptr(), 32 * per, zall[r * per * zt:].data_ptr(), aall[r * per * at:].data_ptr())
    curn_, A, G, cA, cG, D_, _ = nm._draw_arrays(samples)
    phiinv = torch.empty((D, pack.mvar_total), dtype=torch.float64, device="cuda")
    pack.powerlaw_phiinv([s.Ffreqs for s in sigs], A, G, curn.Ffreqs, cA, cG, phiinv.data_ptr())
    out = torch.empty((D, F), dtype=torch.float64, device="cuda")
    pack.nmfp_stage_b(f.data_ptr(), F, zall.data_ptr(), aall.data_ptr(), per, phiinv.data_ptr(), D, out.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), want.view(torch.int64))
    # a single rank: calculate_nmfp_2d is the plain call
    assert torch.equal(nm.calculate_nmfp_2d(f, samples, *mats).view(torch.int64), want.view(torch.int64))
    # an odd number of tiles per block is refused when there are several blocks
    with pytest.raises(fastfp_b200._cabi.FastFpError):
        pack.nmfp_stage_b(f.data_ptr(), F, zall.data_ptr(), aall.data_ptr(), 3, phiinv.data_ptr(), D, out.data_ptr())
