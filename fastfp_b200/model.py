"""Model set-up helpers of the noise-marginalised driver -- the counterpart of the functions the reference
keeps inside ``examples/run_nmfp.py`` (``create_freqarray`` ``:27-35``, ``create_quantization_array`` ``:38-57``,
``ecorr_weights_by_backend`` ``:60-70``, ``setup_fp_model`` ``:73-171``), importable here so a caller script does not
have to carry them. They only need duck-typed pulsars (``.name``, ``.toas``, ``.Mmat``, ``.backend_flags``):
``enterprise`` is not required.

Beyond the reference: :func:`epochs_of` returns the TOA index groups behind the ECORR weights (the reference computes
them and throws them away), from which :func:`ecorr_basis_by_backend` builds the GP-ECORR basis columns that
``enterprise``'s ``white_noise_block(gp_ecorr=True, select="backend")`` would supply, and :func:`kernel_ecorr_blocks`
the block-diagonal ``N`` (:class:`fastfp_b200.BlockNvec`) of the kernel-ECORR alternative (the reference's to-do,
``fastfp/utils.py:29-31``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .blockn import BlockNvec
from .nmfp import NMFP, CURN_container, GPEcorr_container, RN_container


def get_tspan(psrs) -> float:
    """Total time span of the array (``enterprise_extensions.model_utils.get_tspan``; used at
    ``examples/run_nmfp.py:233`` and ``fastfp/utils.py:145``)."""
    return float(max(np.max(p.toas) for p in psrs) - min(np.min(p.toas) for p in psrs))


def create_freqarray(Tspan: float, ncomps: int = 30) -> np.ndarray:
    """``repeat(k / Tspan, 2)``, ``k = 1..ncomps`` (``run_nmfp.py:27-35``)."""
    return np.repeat(1.0 * np.arange(1, ncomps + 1) / Tspan, 2)


def epochs_of(toas, dt: float = 1.0, nmin: int = 2) -> List[np.ndarray]:
    """Observing epochs: walking through the TOAs in time order, a TOA closer than ``dt`` seconds to the FIRST TOA
    of the current bucket joins it, otherwise it opens a new bucket; buckets with fewer than ``nmin`` TOAs are
    dropped (the bucketing of ``run_nmfp.py:42-54``). Returns the index arrays (into ``toas``) of the kept buckets."""
    toas = np.asarray(toas, dtype=np.float64)
    if toas.size == 0:
        return []
    order = np.argsort(toas, kind="stable")
    ts = toas[order]
    starts = [0]
    ref = ts[0]
    for i in range(1, ts.size):
        if not (ts[i] - ref < dt):
            starts.append(i)
            ref = ts[i]
    starts.append(ts.size)
    return [order[a:b] for a, b in zip(starts[:-1], starts[1:]) if b - a >= nmin]


def create_quantization_array(toas, dt: float = 1.0, nmin: int = 2) -> np.ndarray:
    """Per-epoch ECORR weights, hard-wired to 1.0 like the reference (``run_nmfp.py:38-57``): one entry per epoch
    with at least ``nmin`` TOAs."""
    return np.ones(len(epochs_of(toas, dt, nmin)))


def ecorr_weights_by_backend(psr) -> List[np.ndarray]:
    """The weights split by receiver backend, in ``np.unique(psr.backend_flags)`` order (``run_nmfp.py:60-70``) --
    the order :class:`GPEcorr_container` looks the ``log10_ecorr`` values up in (``fastfp/nmfp.py:444-450``)."""
    flags = np.asarray(psr.backend_flags)
    toas = np.asarray(psr.toas, dtype=np.float64)
    return [create_quantization_array(toas[np.nonzero(flags == val)[0]]) for val in np.unique(flags)]


def ecorr_basis_by_backend(psr) -> np.ndarray:
    """Epoch-indicator columns ``U`` (``n x n_epoch``), backend by backend in ``np.unique`` order and epoch by epoch
    in time order inside a backend: the GP-ECORR block of ``T = [timing model | U | Fourier]`` whose prior variances
    :class:`GPEcorr_container` provides. (In the reference these columns come out of ``pta.get_basis``.)"""
    flags = np.asarray(psr.backend_flags)
    toas = np.asarray(psr.toas, dtype=np.float64)
    cols = []
    for val in np.unique(flags):
        idx = np.nonzero(flags == val)[0]
        for ep in epochs_of(toas[idx]):
            u = np.zeros(toas.size)
            u[idx[ep]] = 1.0
            cols.append(u)
    return np.stack(cols, axis=1) if cols else np.zeros((toas.size, 0))


def kernel_ecorr_blocks(psr, Nvec, noise, require_contiguous: bool = True) -> BlockNvec:
    """The same ECORR model as a block-diagonal ``N = diag(Nvec) + sum_e j_e 1_e 1_e^T`` with
    ``j_e = 10^(2 log10_ecorr(backend of e))`` -- ``enterprise``'s ``EcorrKernelNoise`` (the reference's to-do,
    ``README.md:22``). Epochs must be contiguous index ranges (TOAs sorted in time, one backend at a time per epoch),
    which is how ``enterprise`` sorts pulsars for the kernel representation."""
    flags = np.asarray(psr.backend_flags)
    toas = np.asarray(psr.toas, dtype=np.float64)
    slices, jvec = [], []
    for val in np.unique(flags):
        idx = np.nonzero(flags == val)[0]
        j = 10.0 ** (2.0 * float(noise["_".join([psr.name, "basis", "ecorr", str(val), "log10_ecorr"])]))
        for ep in epochs_of(toas[idx]):
            rows = np.sort(idx[ep])
            if require_contiguous and not np.array_equal(rows, np.arange(rows[0], rows[0] + rows.size)):
                raise ValueError(f"{psr.name}: an ECORR epoch is not a contiguous range of TOA indices; sort the "
                                 "TOAs in time (per backend) before using the kernel representation")
            slices.append(slice(int(rows[0]), int(rows[0] + rows.size)))
            jvec.append(j)
    order = np.argsort([s.start for s in slices]) if slices else []
    return BlockNvec(np.asarray(Nvec, dtype=np.float64), [slices[i] for i in order],
                     np.asarray([jvec[i] for i in order], dtype=np.float64))


def setup_fp_model(psrs, noise, Tspan: Optional[float] = None, add_ecorr: bool = False, nrncomps: int = 30,
                   add_curn: bool = False, ngwbcomps: int = 5, device=None) -> NMFP:
    """The :class:`NMFP` object and its red-noise containers (``run_nmfp.py:73-171``; same arguments and defaults).
    ``Tspan=None`` gives every pulsar its own Fourier grid from its own span; a value sets one common grid. The
    common process, when requested, always lives on the array-wide grid with ``ngwbcomps`` components."""
    curn = CURN_container(create_freqarray(get_tspan(psrs), ncomps=ngwbcomps)) if add_curn else None
    rn_objs = []
    for psr in psrs:
        span = Tspan if Tspan else float(np.max(psr.toas) - np.min(psr.toas))
        ecorr = GPEcorr_container(psr, ecorr_weights_by_backend(psr), fix_wn_vals=noise) if add_ecorr else None
        rn_objs.append(RN_container(psr, Ffreqs=create_freqarray(span, ncomps=nrncomps), gp_ecorr=add_ecorr,
                                    ecorr_container=ecorr, add_curn=add_curn, curn_container=curn))
    return NMFP(psrs, rn_objs, device=device)


def param_names(psrs: Sequence, inc_cp: bool) -> List[str]:
    """Chain column order of a PTA built by ``initialize_pta`` with constant white noise: per pulsar
    ``red_noise_gamma``, ``red_noise_log10_A`` (``enterprise`` sorts parameter names), then the common process
    -- what ``pta.params`` yields and ``map_params`` (``run_nmfp.py:174-186``) walks through."""
    names = [f"{psr.name}_red_noise_{k}" for psr in psrs for k in ("gamma", "log10_A")]
    return names + (["gw_gamma", "gw_log10_A"] if inc_cp else [])
