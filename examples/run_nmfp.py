# -*- coding: utf-8 -*-
"""Noise-marginalised Fp with the B200 engine: the counterpart of the reference's
``examples/run_nmfp.py`` (output: ``res/<savefile>.npy`` holding the ``(nsamples, ncwfreqs)`` array).

``--synthetic P N`` uses seeded synthetic pulsars and stand-in MCMC draws; with real data the PTA
and the containers are built as in the reference script (``setup_fp_model`` there works unchanged
with the classes imported from ``fastfp_b200``). All draws go to the GPU in one call: the
``batch_size`` loop of the reference (``run_nmfp.py:256-270``) is not needed.
"""
import argparse
import logging
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastfp_b200 import NMFP, CURN_container, RN_container, chains, get_mats_nmfp, vmap  # noqa: E402


def create_freqarray(Tspan, ncomps=30):
    return np.repeat(1.0 * np.arange(1, ncomps + 1) / Tspan, 2)


def main(synthetic, savefile="nmfp_out", inc_cp=True, nrncomps=30, ngwbcomps=30, ncwfreqs=100, nsamples=1000,
         chainfile=None, batch_size=None):
    logging.basicConfig(format="%(levelname)s: %(message)s", level=logging.INFO)
    logger = logging.getLogger(__name__)
    from fastfp_b200 import synth

    pta = synth.make_pta(synthetic[0], synthetic[1], ncomps=nrncomps, inc_cp=inc_cp)
    psrs, noise, Tspan = pta.psrs, pta.noise, pta.Tspan
    Ffreqs_rn = create_freqarray(Tspan, ncomps=nrncomps)
    if inc_cp:
        curn_obj = CURN_container(create_freqarray(Tspan, ncomps=ngwbcomps))
        rn_objs = [RN_container(psr, Ffreqs=Ffreqs_rn, add_curn=True, curn_container=curn_obj) for psr in psrs]
    else:
        rn_objs = [RN_container(psr, Ffreqs=Ffreqs_rn) for psr in psrs]
    nmfp = NMFP(psrs, rn_objs)

    t_start = time.perf_counter()
    TNTs, Nvecs, Ts = get_mats_nmfp(pta, noise)
    logger.info(f"Precompute matrix wall time: {time.perf_counter() - t_start:.2f} s")

    freqs = np.arange(1, ncwfreqs + 1) / Tspan
    # parameter order of the chain columns = pta.params order in the reference (run_nmfp.py:174-186)
    param_names = [f"{psr.name}_red_noise_{k}" for psr in psrs for k in ("gamma", "log10_A")] + \
                  (["gw_gamma", "gw_log10_A"] if inc_cp else [])
    if chainfile is None:  # stand-in MCMC chain, written in the PTMCMC text layout and read back
        os.makedirs("res", exist_ok=True)
        chainfile = f"res/{savefile}_chain_1.txt"
        chains.write_chain(chainfile, synth.draw_samples(pta, 2 * nsamples), param_names)
    samples, _ = chains.draws_from_chain(chainfile, param_names, nsamples)  # 25% burn-in, distinct rows

    t_start = time.perf_counter()
    vmap_f = vmap(nmfp, in_axes=(0, None, None, None, None))
    vmap_g = vmap(vmap_f, in_axes=(None, 0, None, None, None))
    if batch_size:  # the reference's draw batches (run_nmfp.py:256-270); only bounds the host-side output here
        nmfp_vals = np.vstack([np.asarray(vmap_g(freqs, part, Nvecs, Ts, TNTs))
                               for part in chains.draw_batches(samples, batch_size)])
    else:
        nmfp_vals = np.asarray(vmap_g(freqs, samples, Nvecs, Ts, TNTs))
    logger.info(f"Noise marginalized Fp-statistic wall time: {time.perf_counter() - t_start:.2f} s")

    os.makedirs("res", exist_ok=True)
    with open(f"res/{savefile}.npy", "wb") as f:
        np.save(f, nmfp_vals)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--synthetic", nargs=2, type=int, metavar=("P", "NTOA"), default=[10, 2000])
    parser.add_argument("--savefile", type=str, default="nmfp_out")
    parser.add_argument("--inc_cp", action="store_true", help="include CURN process")
    parser.add_argument("--nrncomps", type=int, default=30)
    parser.add_argument("--ngwbcomps", type=int, default=30)
    parser.add_argument("--ncwfreqs", type=int, default=100)
    parser.add_argument("--nsamples", type=int, default=1000)
    parser.add_argument("--chainfile", type=str, default=None, help="PTMCMC text chain (last 4 columns = sampler metadata)")
    parser.add_argument("--batch_size", type=int, default=None)
    main(**vars(parser.parse_args()))
