"""Host-side model set-up helpers (fastfp_b200/model.py): the functions the reference keeps in examples/run_nmfp.py
(create_quantization_array :38-57, ecorr_weights_by_backend :60-70, setup_fp_model :73-171), checked on the CPU."""
import numpy as np
import pytest

from conftest import Psr
from fastfp_b200 import model, synth
from oracle import fp_oracle as o


def _reference_bucketing(toas, dt=1, nmin=2):
    """the bucketing rule of run_nmfp.py:42-54 restated literally (sorted walk, first TOA of a bucket is its reference)"""
    isort = np.argsort(toas)
    ref, ind = [toas[isort[0]]], [[isort[0]]]
    for i in isort[1:]:
        if toas[i] - ref[-1] < dt:
            ind[-1].append(i)
        else:
            ref.append(toas[i])
            ind.append([i])
    return [b for b in ind if len(b) >= nmin]


def test_epochs_follow_the_reference_bucketing_rule():
    rng = np.random.default_rng(3)
    t0 = np.sort(rng.uniform(0, 1e6, 40))
    toas = np.concatenate([t + np.arange(k) * 0.3 for t, k in zip(t0, rng.integers(1, 6, 40))])
    toas = toas[rng.permutation(toas.size)]  # unsorted input
    want = _reference_bucketing(toas)
    got = model.epochs_of(toas)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(np.sort(g), np.sort(np.asarray(w)))
    w = model.create_quantization_array(toas)
    assert w.shape == (len(want),) and np.all(w == 1.0)
    # a chain of TOAs 0.6 s apart: buckets are anchored at their FIRST member, so it splits every second TOA
    chain = np.arange(7) * 0.6
    assert [len(e) for e in model.epochs_of(chain)] == [2, 2, 2]
    assert model.epochs_of(np.array([1.0, 5.0, 9.0])) == [] and model.epochs_of(np.zeros(0)) == []


def test_weights_basis_and_kernel_blocks_agree():
    pta = synth.make_pta(2, [203, 160], n_tm=[4, 5], ncomps=6, seed=3, epoch=4, nbackends=3)
    q = pta.psrs[0]
    w = model.ecorr_weights_by_backend(q)
    assert len(w) == 3 and [f for f in np.unique(q.backend_flags)] == ["be0", "be1", "be2"]
    U = model.ecorr_basis_by_backend(q)
    assert U.shape == (203, sum(len(x) for x in w))
    assert set(np.unique(U)) == {0.0, 1.0} and np.all(U.sum(1) <= 1) and np.all(U.sum(0) >= 2)
    # every column lives inside one backend, columns ordered backend by backend
    owner = [np.unique(q.backend_flags[U[:, e] > 0]) for e in range(U.shape[1])]
    assert all(len(o_) == 1 for o_ in owner)
    assert [o_[0] for o_ in owner] == sorted(o_[0] for o_ in owner)
    Nvecs, Ts, TNTs, phis = synth.with_ecorr(pta)
    B = model.kernel_ecorr_blocks(q, pta.Nvecs[0], pta.noise)
    ntm = pta.n_tm[0]
    jv = phis[0][ntm:ntm + U.shape[1]]
    np.testing.assert_allclose(B.dense(), np.diag(pta.Nvecs[0]) + (U * jv) @ U.T, rtol=1e-15, atol=0)
    assert Ts[0].shape[1] == pta.Ts[0].shape[1] + U.shape[1]
    np.testing.assert_array_equal(Ts[0][:, ntm:ntm + U.shape[1]], U)


@pytest.mark.parametrize("add_ecorr", [False, True])
@pytest.mark.parametrize("add_curn", [False, True])
@pytest.mark.parametrize("common_span", [False, True])
def test_setup_fp_model_selects_the_reference_layouts(add_ecorr, add_curn, common_span):
    pta = synth.make_pta(3, [120, 160, 140], n_tm=[3, 4, 5], ncomps=7, seed=5, epoch=4)
    synth.with_ecorr(pta)  # fills the ecorr keys of pta.noise
    nm = model.setup_fp_model(pta.psrs, pta.noise, Tspan=pta.Tspan if common_span else None, add_ecorr=add_ecorr,
                              nrncomps=7, add_curn=add_curn, ngwbcomps=4)
    name = "get_phi_tm" + ("_ecorr" if add_ecorr else "") + "_rn" + ("_curn" if add_curn else "")
    pars = dict(pta.noise)
    for p, (sig, q) in enumerate(zip(nm.rn_sigs, pta.psrs)):
        assert sig.phi_fn.__name__ == name
        span = pta.Tspan if common_span else q.toas.max() - q.toas.min()
        np.testing.assert_array_equal(sig.Ffreqs, np.repeat(np.arange(1, 8) / span, 2))
        ec = None
        if add_ecorr:
            vals = [pta.noise[f"{q.name}_basis_ecorr_{b}_log10_ecorr"] for b in np.unique(q.backend_flags)]
            ec = o.ecorr_phi(model.ecorr_weights_by_backend(q), vals)
        want = o.get_phi(pars, q.name, pta.n_tm[p], sig.Ffreqs, add_curn=add_curn,
                         curn_Ffreqs=np.repeat(np.arange(1, 5) / model.get_tspan(pta.psrs), 2), ecorr_phi_fixed=ec)
        np.testing.assert_array_equal(sig.update_phi(pars), want)
    if add_curn:
        assert all(s.curn_container is nm.rn_sigs[0].curn_container for s in nm.rn_sigs)


def test_param_names_and_tspan():
    psrs = [Psr(np.array([0.0, 10.0]), np.zeros(2), name="A"), Psr(np.array([5.0, 30.0]), np.zeros(2), name="B")]
    assert model.get_tspan(psrs) == 30.0
    assert model.param_names(psrs, True) == ["A_red_noise_gamma", "A_red_noise_log10_A", "B_red_noise_gamma",
                                             "B_red_noise_log10_A", "gw_gamma", "gw_log10_A"]
    assert model.param_names(psrs, False)[-1] == "B_red_noise_log10_A"
