"""EXTENDED-PRECISION TRUTH -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The same quantities as ``oracle/fp_oracle.py`` evaluated in x87 ``np.longdouble`` (64-bit
mantissa, eps ~ 1.1e-19), plus a first-order *conditioning* figure for every output.

Why it exists: near a red-noise Fourier frequency the Earth-term basis
lies almost inside span(T) and ``(x|y) = x^T N^-1 y - (T^T N^-1 x)^T Sigma^-1 (T^T N^-1 y)``
is the difference of two numbers up to 1e6-1e10 times larger than the result. There the
reference's own float64 output is only defined to ``eps * kappa``; two correct float64
implementations that differ in summation order differ by that much. Parity tests therefore
use ``|cuda - oracle| <= 1e-10*|oracle| + c*eps*cond`` with ``cond`` from this module, and
additionally check that the CUDA path is no further from this truth than the oracle is.

Definition of "truth": the float64 inputs are exact, and the float64-rounded phase
``fl(fl(2*pi*f)*t)`` of reference ``fastfp/fastfp.py:78-79`` is taken as exact input (the
reference defines the phase by that float64 expression); everything after that -- sin/cos,
N^-1 scaling, the T^T N^-1 x products, the Sigma solve, the 2x2 solve -- is done in longdouble.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def _chol_ld(S):
    """Lower Cholesky factor in longdouble (no LAPACK for this dtype)."""
    A = np.array(S, dtype=LD)
    m = A.shape[0]
    L = np.zeros_like(A)
    for j in range(m):
        d = A[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(d)
        if j + 1 < m:
            L[j + 1 :, j] = (A[j + 1 :, j] - L[j + 1 :, :j] @ L[j, :j]) / L[j, j]
    return L


def _fwd_ld(L, B):
    """Solve L X = B (lower triangular), longdouble, B: (m, k)."""
    m = L.shape[0]
    X = np.array(B, dtype=LD)
    for j in range(m):
        X[j] = (X[j] - L[j, :j] @ X[:j]) / L[j, j]
    return X


def fp_sweep_truth(freqs, toas, residuals, Nvecs, Ts, sigmas, chunk=128):
    """Truth for ``fp_sweep``. Returns ``(terms (P,F) longdouble, cond (P,F) float64)``.

    ``cond[p, f]`` bounds (to first order, in units of the relative rounding error committed
    in each of the two cancelling parts of every inner product) the absolute change of the
    per-pulsar term ``0.5 * N^T M^-1 N``:
    ``|x|^T A + 0.5 |x|^T B |x|`` with ``x = M^-1 N`` and ``A_k``/``B_kl`` the sums of the
    magnitudes of the two parts of ``N_k``/``M_kl``."""
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    F, P = freqs.shape[0], len(toas)
    terms = np.zeros((P, F), dtype=LD)
    cond = np.zeros((P, F))
    for p in range(P):
        toa = np.asarray(toas[p], dtype=np.float64)
        ninv = LD(1) / np.asarray(Nvecs[p], dtype=LD)
        T = np.asarray(Ts[p], dtype=LD)
        r = np.asarray(residuals[p], dtype=LD)
        L = _chol_ld(sigmas[p])
        G = _fwd_ld(L, (T * ninv[:, None]).T)  # (m, n): L^-1 T^T N^-1
        ur = G @ r
        rn = r * ninv
        for lo in range(0, F, chunk):
            f = freqs[lo : lo + chunk]
            # float64 phase exactly as the reference forms it: ((2*pi)*f)*t
            ph = ((2 * np.pi * f)[:, None] * toa[None, :]).astype(LD)
            S, C = np.sin(ph), np.cos(ph)
            US, UC = S @ G.T, C @ G.T  # (F, m)
            Sn, Cn = S * ninv, C * ninv
            sNs, sNc, cNc = (S * Sn).sum(1), (S * Cn).sum(1), (C * Cn).sum(1)
            sNr, cNr = S @ rn, C @ rn
            bss, bsc, bcc = (US * US).sum(1), (US * UC).sum(1), (UC * UC).sum(1)
            bsr, bcr = US @ ur, UC @ ur
            Mss, Msc, Mcc = sNs - bss, sNc - bsc, cNc - bcc
            Ns, Nc = sNr - bsr, cNr - bcr
            det = Mss * Mcc - Msc * Msc
            x0 = (Mcc * Ns - Msc * Nc) / det
            x1 = (Mss * Nc - Msc * Ns) / det
            terms[p, lo : lo + chunk] = LD(0.5) * (Ns * x0 + Nc * x1)
            A0 = np.abs(sNr) + np.abs(bsr)
            A1 = np.abs(cNr) + np.abs(bcr)
            B00, B01, B11 = sNs + bss, np.abs(sNc) + np.abs(bsc), cNc + bcc
            ax0, ax1 = np.abs(x0), np.abs(x1)
            c = ax0 * A0 + ax1 * A1 + LD(0.5) * (ax0 * ax0 * B00 + 2 * ax0 * ax1 * B01 + ax1 * ax1 * B11)
            cond[p, lo : lo + chunk] = c.astype(np.float64)
    return terms, cond


def get_xCy_truth(Nvec, T, sigma, x, y):
    """Truth for one ``get_xCy`` (reference ``fastfp/utils.py:49-54``); returns
    ``(value longdouble, cond float64)`` with ``cond = |x N^-1 y| + |second term|``."""
    ninv = LD(1) / np.asarray(Nvec, dtype=LD)
    T = np.asarray(T, dtype=LD)
    x = np.asarray(x, dtype=LD)
    y = np.asarray(y, dtype=LD)
    L = _chol_ld(sigma)
    ux = _fwd_ld(L, (T.T @ (x * ninv))[:, None])[:, 0]
    uy = _fwd_ld(L, (T.T @ (y * ninv))[:, None])[:, 0]
    a = (x * ninv * y).sum()
    b = ux @ uy
    return a - b, float(np.abs(a) + np.abs(b))


def fp_sweep_truth_blockn(freqs, toas, residuals, blocks, Ts, phiinvs=None, chunk=64, sigmas=None):
    """Truth for a BLOCK-DIAGONAL white-noise matrix ``N = diag(nvec) + sum_e j_e 1_e 1_e^T`` (kernel ECORR).

    The reference has no implementation of this case (``fastfp/utils.py:29-31``); it is mathematically the
    GP-basis model ``C = diag(nvec) + U J U^T + T Phi T^T`` the reference does implement (epoch-indicator
    columns appended to ``T``, ``fastfp/nmfp.py:277-282``), whose extended-precision evaluation
    (:func:`fp_sweep_truth` on the widened basis) costs O((m + n_epoch)^3) longdouble operations without
    LAPACK -- minutes per pulsar at 2500 epochs. This function evaluates the SAME quantity with ``N^-1``
    applied by Sherman-Morrison in longdouble, ``Sigma = T^T N^-1 T + diag(phiinv)`` formed in longdouble;
    ``tests/test_oracle_golden.py`` pins it against :func:`fp_sweep_truth` on the widened basis at a size
    where both run. ``blocks[p]`` is ``(nvec, [(start, stop), ...], jvec)``. With ``sigmas`` given, those
    float64 matrices are taken as exact inputs (what the engine is handed), like :func:`fp_sweep_truth` does;
    otherwise ``Sigma`` is formed here from ``phiinvs``. Returns ``(terms, cond)`` like :func:`fp_sweep_truth`."""
    freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    F, P = freqs.shape[0], len(toas)
    terms = np.zeros((P, F), dtype=LD)
    cond = np.zeros((P, F))
    for p in range(P):
        toa = np.asarray(toas[p], dtype=np.float64)
        nvec, slices, jvec = blocks[p]
        ninv = LD(1) / np.asarray(nvec, dtype=LD)
        T = np.asarray(Ts[p], dtype=LD)
        r = np.asarray(residuals[p], dtype=LD)
        n = toa.shape[0]
        # epoch membership: eid[i] = epoch of TOA i or -1; beta_e = j_e / (1 + j_e sum_e 1/nvec)
        eid = np.full(n, -1, dtype=np.int64)
        for e, (a, b) in enumerate(slices):
            eid[a:b] = e
        ne = len(slices)
        member = eid >= 0
        ssum = np.zeros(ne, dtype=LD)
        np.add.at(ssum, eid[member], ninv[member])
        jv = np.asarray(jvec, dtype=LD)
        beta = jv / (LD(1) + jv * ssum)

        def nsolve(X):  # N^-1 X, X: (n,) or (n, k)
            Y = X * (ninv if X.ndim == 1 else ninv[:, None])
            acc = np.zeros((ne,) + Y.shape[1:], dtype=LD)
            np.add.at(acc, eid[member], Y[member])
            corr = (beta if X.ndim == 1 else beta[:, None]) * acc
            Y = Y.copy()
            Y[member] -= (ninv[member] if X.ndim == 1 else ninv[member][:, None]) * corr[eid[member]]
            return Y

        NT = nsolve(T)                       # N^-1 T   (n, m)
        if sigmas is not None:
            Sigma = np.asarray(sigmas[p], dtype=LD)
        else:
            Sigma = T.T @ NT + np.diag(np.asarray(phiinvs[p], dtype=LD))
        L = _chol_ld(Sigma)
        G = _fwd_ld(L, NT.T)                 # L^-1 T^T N^-1   (m, n)
        rn = nsolve(r)
        ur = G @ r
        for lo in range(0, F, chunk):
            f = freqs[lo : lo + chunk]
            ph = ((2 * np.pi * f)[:, None] * toa[None, :]).astype(LD)
            S, C = np.sin(ph), np.cos(ph)
            Sn, Cn = nsolve(S.T).T, nsolve(C.T).T
            US, UC = S @ G.T, C @ G.T
            sNs, sNc, cNc = (S * Sn).sum(1), (S * Cn).sum(1), (C * Cn).sum(1)
            sNr, cNr = S @ rn, C @ rn
            bss, bsc, bcc = (US * US).sum(1), (US * UC).sum(1), (UC * UC).sum(1)
            bsr, bcr = US @ ur, UC @ ur
            Mss, Msc, Mcc = sNs - bss, sNc - bsc, cNc - bcc
            Ns, Nc = sNr - bsr, cNr - bcr
            det = Mss * Mcc - Msc * Msc
            x0 = (Mcc * Ns - Msc * Nc) / det
            x1 = (Mss * Nc - Msc * Ns) / det
            terms[p, lo : lo + chunk] = LD(0.5) * (Ns * x0 + Nc * x1)
            # conditioning: the diagonal-N part and the two subtracted parts (epoch correction, Woodbury)
            dS, dC = S * ninv, C * ninv
            pss, psc, pcc = (S * dS).sum(1), np.abs(S * dC).sum(1), (C * dC).sum(1)
            A0 = np.abs(S * (r * ninv)).sum(1) + np.abs(bsr)
            A1 = np.abs(C * (r * ninv)).sum(1) + np.abs(bcr)
            B00, B01, B11 = 2 * pss - sNs + bss, psc + np.abs(bsc), 2 * pcc - cNc + bcc
            ax0, ax1 = np.abs(x0), np.abs(x1)
            c = ax0 * A0 + ax1 * A1 + LD(0.5) * (ax0 * ax0 * B00 + 2 * ax0 * ax1 * B01 + ax1 * ax1 * B11)
            cond[p, lo : lo + chunk] = c.astype(np.float64)
    return terms, cond
