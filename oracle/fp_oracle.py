"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 NumPy restatement of the reference's Fp hot path (gabefreedman/fastfp @ 74b0ef8),
written from the formulas, operation for operation, so that parity tests can compare the
CUDA path against "what the reference computes". Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module.

Pinning: the reference ships no tests or golden vectors and JAX is not
installable in this image, so the oracle is pinned against **the reference's own source
files executed here under a NumPy-backed ``jax`` shim** (``tests/golden/make_golden.py``;
fixtures ``tests/golden/*.npz``). What that pins: formulas, argument order, operation order,
layouts. What it cannot pin: XLA's primitive rounding (its ``sin``/``cos``/``dot``/LU), which
is not reproducible without XLA.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np

# reference fastfp/constants.py:7-9 (scipy.constants.Julian_year = 365.25 d)
yr = 31557600.0
fyr = 1.0 / yr


# --------------------------------------------------------------------------------------
# get_xCy                                                        reference fastfp/utils.py:26-54
# --------------------------------------------------------------------------------------
def get_xCy(Nvec, T, sigma, x, y):
    """x^T C^-1 y with C = N + T B T^T, diagonal N (reference ``fastfp/utils.py:49-54``)."""
    Nx = np.array(x / Nvec)  # utils.py:49
    Ny = np.array(y / Nvec)  # utils.py:50
    TNx = np.dot(T.T, Nx)  # utils.py:51
    TNy = np.dot(T.T, Ny)  # utils.py:52
    xNy = np.dot(x.T, Ny)  # utils.py:53
    return xNy - TNx @ np.linalg.solve(sigma, TNy)  # utils.py:54


# --------------------------------------------------------------------------------------
# FastFp.calculate_Fp                                            reference fastfp/fastfp.py:51-92
# --------------------------------------------------------------------------------------
def calculate_Fp(fgw, toas, residuals, Nvecs, Ts, sigmas):
    """One GW frequency, literal per-pulsar loop (reference ``fastfp/fastfp.py:69-92``).

    ``toas``/``residuals`` are the per-pulsar lists the reference stores on ``self``
    (``fastfp.py:44-45``)."""
    fstat = 0  # fastfp.py:71
    for Nvec, T, sigma, toa, resid in zip(Nvecs, Ts, sigmas, toas, residuals):  # :72-74
        ntoa = toa.shape[0]
        A = np.zeros((2, ntoa))
        # phase is ((2*pi)*fgw)*toa, one rounding per multiply                  # :78-79
        A[0, :] = 1 / fgw ** (1 / 3) * np.sin(2 * np.pi * fgw * toa)
        A[1, :] = 1 / fgw ** (1 / 3) * np.cos(2 * np.pi * fgw * toa)
        ip1 = get_xCy(Nvec, T, sigma, A[0, :], resid)  # :81
        ip2 = get_xCy(Nvec, T, sigma, A[1, :], resid)  # :82
        N = np.array([ip1, ip2])  # :83
        M = np.zeros((2, 2))
        M[0, 0] = get_xCy(Nvec, T, sigma, A[0, :], A[0, :])  # :85
        M[0, 1] = get_xCy(Nvec, T, sigma, A[0, :], A[1, :])  # :86
        M[1, 0] = get_xCy(Nvec, T, sigma, A[1, :], A[0, :])  # :87
        M[1, 1] = get_xCy(Nvec, T, sigma, A[1, :], A[1, :])  # :88
        fstat += 0.5 * np.dot(N, np.linalg.solve(M, N))  # :90
    return fstat


def _solve2x2_batched(M, N):
    """Batched general 2x2 solve with partial pivoting (what ``jnp.linalg.solve`` does for
    ``fastfp.py:90``), vectorised over a leading frequency axis. M: (F,2,2), N: (F,2)."""
    a, b, c, d = M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1]
    n0, n1 = N[:, 0], N[:, 1]
    swap = np.abs(c) > np.abs(a)
    a_, b_, c_, d_ = np.where(swap, c, a), np.where(swap, d, b), np.where(swap, a, c), np.where(swap, b, d)
    m0, m1 = np.where(swap, n1, n0), np.where(swap, n0, n1)
    with np.errstate(all="ignore"):
        l = c_ / a_
        u = d_ - l * b_
        x1 = (m1 - l * m0) / u
        x0 = (m0 - b_ * x1) / a_
    return np.stack((x0, x1), axis=1)


def fp_sweep(freqs, toas, residuals, Nvecs, Ts, sigmas, chunk=256, per_pulsar=False):
    """``jax.vmap(calculate_Fp, in_axes=(0, None, None, None))`` (reference
    ``examples/run_fp.py:63-64``) in batched-over-frequency form: the same formulas as
    :func:`calculate_Fp`, with a leading ``(F_chunk,)`` axis on ``A`` the way ``vmap``
    materialises it, the three ``T^T N^-1 x`` products as GEMMs and one LU of ``sigma`` per
    pulsar with batched right-hand sides (the sharing XLA's CSE gives the six ``get_xCy``
    calls). Chunked over frequency so the ``(F, n)`` temporaries fit in RAM.

    Returns ``(F,)``; with ``per_pulsar=True`` returns the ``(P, F)`` per-pulsar terms."""
    import scipy.linalg as sla

    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    F = freqs.shape[0]
    P = len(toas)
    terms = np.zeros((P, F))
    lus = [sla.lu_factor(s) for s in sigmas]
    for lo in range(0, F, chunk):
        f = freqs[lo : lo + chunk]
        pref = 1 / f ** (1 / 3)
        w = (2 * np.pi * f)[:, None]
        for p, (Nvec, T, lu, toa, resid) in enumerate(zip(Nvecs, Ts, lus, toas, residuals)):
            ph = w * toa[None, :]
            S = pref[:, None] * np.sin(ph)
            C = pref[:, None] * np.cos(ph)
            NS, NC, Nr = S / Nvec, C / Nvec, resid / Nvec
            TNS, TNC, TNr = NS @ T, NC @ T, T.T @ Nr
            sNr, cNr = S @ Nr, C @ Nr
            sNs = np.einsum("fi,fi->f", S, NS)
            sNc = np.einsum("fi,fi->f", S, NC)
            cNs = np.einsum("fi,fi->f", C, NS)
            cNc = np.einsum("fi,fi->f", C, NC)
            SiS = sla.lu_solve(lu, TNS.T).T
            SiC = sla.lu_solve(lu, TNC.T).T
            Sir = sla.lu_solve(lu, TNr)
            N = np.stack((sNr - TNS @ Sir, cNr - TNC @ Sir), axis=1)
            M = np.empty((f.shape[0], 2, 2))
            M[:, 0, 0] = sNs - np.einsum("fj,fj->f", TNS, SiS)
            M[:, 0, 1] = sNc - np.einsum("fj,fj->f", TNS, SiC)
            M[:, 1, 0] = cNs - np.einsum("fj,fj->f", TNC, SiS)
            M[:, 1, 1] = cNc - np.einsum("fj,fj->f", TNC, SiC)
            x = _solve2x2_batched(M, N)
            terms[p, lo : lo + chunk] = 0.5 * np.einsum("fk,fk->f", N, x)
    if per_pulsar:
        return terms
    out = np.zeros(F)
    for p in range(P):  # sequential pulsar sum starting from 0, fastfp.py:71,90
        out = out + terms[p]
    return out


def fp_sweep_mt(freqs, toas, residuals, Nvecs, Ts, sigmas, workers=None, chunk=128):
    """:func:`fp_sweep` spread over the host cores: the (pulsar, frequency-chunk) pieces are
    independent, so they run on a thread pool (NumPy releases the GIL in sin/cos/GEMM) with BLAS
    limited to one thread per worker. Same arithmetic per piece as :func:`fp_sweep`; used as the
    CPU baseline of bench.py (all host threads), not as a parity reference."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from threadpoolctl import threadpool_limits

    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    F, P = freqs.shape[0], len(toas)
    workers = workers or os.cpu_count() or 1
    terms = np.zeros((P, F))
    tasks = [(p, lo) for lo in range(0, F, chunk) for p in range(P)]

    def run(task):
        p, lo = task
        terms[p, lo : lo + chunk] = fp_sweep(
            freqs[lo : lo + chunk], [toas[p]], [residuals[p]], [Nvecs[p]], [Ts[p]], [sigmas[p]], chunk=chunk
        )

    with threadpool_limits(limits=1):
        with ThreadPoolExecutor(max_workers=min(workers, len(tasks))) as ex:
            list(ex.map(run, tasks))
    out = np.zeros(F)
    for p in range(P):
        out = out + terms[p]
    return out


# --------------------------------------------------------------------------------------
# red-noise containers                                           reference fastfp/nmfp.py:131-477
# --------------------------------------------------------------------------------------
def create_freqarray(psr_toas, ncomps=30):
    """``RN_container._create_freqarray`` (reference ``fastfp/nmfp.py:201-215``)."""
    Tspan = np.max(psr_toas) - np.min(psr_toas)
    f = 1.0 * np.arange(1, ncomps + 1) / Tspan
    return np.repeat(f, 2)


def powerlaw(Ffreqs, log10_A, gamma):
    """``RN_container._powerlaw`` / ``CURN_container._powerlaw`` (reference
    ``fastfp/nmfp.py:226-234`` and ``:373-381``), left-to-right operation order."""
    df = np.diff(np.concatenate((np.array([0]), Ffreqs[::2])))
    return (
        Ffreqs ** (-gamma)
        * (10**log10_A) ** 2
        / 12.0
        / np.pi**2
        * fyr ** (gamma - 3)
        * np.repeat(df, 2)
    )


def ecorr_phi(weights, log10_ecorrs):
    """``GPEcorr_container._init_phi`` (reference ``fastfp/nmfp.py:454-461``)."""
    return np.concatenate([np.asarray(w) * 10 ** (2 * e) for w, e in zip(weights, log10_ecorrs)])


def get_phi(
    pars,
    psr_name,
    n_tm,
    Ffreqs,
    add_curn=False,
    curn_Ffreqs=None,
    ecorr_phi_fixed=None,
):
    """The four selectable phi layouts ``get_phi_tm_rn[_curn]`` / ``get_phi_tm_ecorr_rn[_curn]``
    (reference ``fastfp/nmfp.py:264-292``; selector ``:188-199``):
    ``[ones(n_tm)*1e40 | ecorr phi (fixed) | rn phi (+ curn phi on the leading entries)]``."""
    rn_phi = powerlaw(Ffreqs, pars[f"{psr_name}_red_noise_log10_A"], pars[f"{psr_name}_red_noise_gamma"])
    if add_curn:
        curn_phi = powerlaw(curn_Ffreqs, pars["gw_log10_A"], pars["gw_gamma"])  # nmfp.py:357-358
        rn_phi = rn_phi.copy()
        rn_phi[: curn_phi.shape[0]] += curn_phi  # nmfp.py:275
    tm_phi = np.ones(n_tm) * 1e40  # nmfp.py:267
    if ecorr_phi_fixed is not None:
        return np.concatenate((tm_phi, ecorr_phi_fixed, rn_phi))  # nmfp.py:282
    return np.concatenate((tm_phi, rn_phi))  # nmfp.py:268


def get_phiinv(*a, **k):
    """``RN_container.get_phiinv`` (reference ``fastfp/nmfp.py:305-315``)."""
    return 1.0 / get_phi(*a, **k)


def get_sigmas(pars, TNTs, phi_args):
    """``NMFP._get_sigmas`` (reference ``fastfp/nmfp.py:57-74``). ``phi_args[p]`` holds the
    keyword arguments of :func:`get_phi` for pulsar ``p``."""
    sigmas = []
    for TNT, kw in zip(TNTs, phi_args):
        phiinv = get_phiinv(pars, **kw)
        sigmas.append(TNT + np.diag(phiinv))  # nmfp.py:73
    return sigmas


def calculate_nmfp(fgw, samples, toas, residuals, Nvecs, Ts, TNTs, phi_args):
    """``NMFP.calculate_nmfp`` (reference ``fastfp/nmfp.py:76-119``): ``_get_sigmas`` and then
    the same per-pulsar loop as ``calculate_Fp`` (``:96-119`` is a verbatim copy of
    ``fastfp.py:69-92``)."""
    sigmas = get_sigmas(samples, TNTs, phi_args)  # nmfp.py:93
    return calculate_Fp(fgw, toas, residuals, Nvecs, Ts, sigmas)


def nmfp_sweep(freqs, samples, toas, residuals, Nvecs, Ts, TNTs, phi_args, chunk=256):
    """The double ``vmap`` of ``examples/run_nmfp.py:265-270``: output ``(D, F)``, draw-major.
    ``samples`` is the dict name -> ``(D,)`` array that ``map_params`` builds (``:174-186``)."""
    D = len(next(iter(samples.values())))
    out = np.empty((D, np.atleast_1d(freqs).shape[0]))
    for d in range(D):
        pars = {k: v[d] for k, v in samples.items()}
        sigmas = get_sigmas(pars, TNTs, phi_args)
        out[d] = fp_sweep(freqs, toas, residuals, Nvecs, Ts, sigmas, chunk=chunk)
    return out
