"""The two example drivers (reference examples/run_fp.py, examples/run_nmfp.py) executed
end to end on synthetic inputs, their on-disk outputs (JSON {freq: Fp}; .npy (nsamples, ncwfreqs)) checked against
the oracle / the longdouble truth of the reference formulas."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import EPS
from fastfp_b200 import chains, model, synth
from oracle import fp_oracle as o
from oracle import truth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *map(str, args)], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    return out


def test_run_fp_script_writes_the_reference_json(tmp_path):
    save = tmp_path / "fp"
    _run("run_fp.py", "--synthetic", 3, 300, "--nfreqs", 25, "--save", save)
    with open(str(save) + ".json") as f:
        res = json.load(f)
    freqs = np.linspace(2e-9, 3e-7, 25)  # run_fp.py:59
    assert [float(k) for k in res] == freqs.tolist()
    got = np.array(list(res.values()))
    pta = synth.make_pta(3, 300)
    tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, pta.sigmas)
    tv = tt.sum(0).astype(float)
    assert np.all(np.abs(got - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0))
    want = o.fp_sweep(freqs, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, pta.sigmas)
    well = freqs > 40.0 / pta.Tspan
    assert np.abs(got[well] / want[well] - 1).max() < 1e-10


@pytest.mark.parametrize("flags", [(), ("--inc_cp",), ("--inc_cp", "--inc_ecorr"), ("--inc_cp", "--kernel_ecorr"),
                                   ("--inc_cp", "--batch_size", "4")])
def test_run_nmfp_script_writes_the_reference_array(tmp_path, flags):
    P, n, F, D, nrn, ngw = 3, 400, 12, 6, 8, 5
    _run("run_nmfp.py", "--synthetic", P, n, "--ncwfreqs", F, "--nsamples", D, "--nrncomps", nrn, "--ngwbcomps", ngw,
         "--seed", 1, "--outdir", tmp_path, "--save", "out", *flags)
    got = np.load(tmp_path / "out.npy")
    rows = np.load(tmp_path / "out_rows.npy")
    assert got.shape == (D, F) and rows.shape == (D,) and len(set(rows.tolist())) == D
    inc_cp, ecorr = "--inc_cp" in flags, "--inc_ecorr" in flags or "--kernel_ecorr" in flags
    pta = synth.make_pta(P, n, ncomps=nrn, inc_cp=inc_cp, epoch=4 if ecorr else 0)
    names = model.param_names(pta.psrs, inc_cp)
    chain = np.loadtxt(tmp_path / "out_chain_1.txt")
    assert rows.min() >= int(0.25 * chain.shape[0])  # burn-in respected (run_nmfp.py:221-222)
    samples = chains.map_params(names, chain[rows, :len(names)].T)
    Ts, Nvecs, TNTs = pta.Ts, pta.Nvecs, pta.TNTs
    ec = [None] * P
    if ecorr:  # the oracle always evaluates the GP form the reference implements; --kernel_ecorr must agree with it
        Nvecs, Ts, TNTs, _ = synth.with_ecorr(pta)
        for p, q in enumerate(pta.psrs):
            vals = [pta.noise[f"{q.name}_basis_ecorr_{b}_log10_ecorr"] for b in np.unique(q.backend_flags)]
            ec[p] = o.ecorr_phi(model.ecorr_weights_by_backend(q), vals)
    Tspan = model.get_tspan(pta.psrs)
    Ffreqs = np.repeat(np.arange(1, nrn + 1) / Tspan, 2)
    phi_args = [dict(psr_name=q.name, n_tm=pta.n_tm[p], Ffreqs=Ffreqs, add_curn=inc_cp,
                     curn_Ffreqs=np.repeat(np.arange(1, ngw + 1) / Tspan, 2), ecorr_phi_fixed=ec[p])
                for p, q in enumerate(pta.psrs)]
    freqs = np.arange(1, F + 1) / Tspan
    for d in (0, D - 1):
        pars = {k: v[d] for k, v in samples.items()}
        sig = o.get_sigmas(pars, TNTs, phi_args)
        tt, cond = truth.fp_sweep_truth(freqs, pta.toas, pta.residuals, Nvecs, Ts, sig)
        tv = tt.sum(0).astype(float)
        assert np.all(np.abs(got[d] - tv) <= 1e-10 * np.abs(tv) + 256 * EPS * cond.sum(0)), (flags, d)
