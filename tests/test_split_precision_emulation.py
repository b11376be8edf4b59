"""CPU emulation of the round-2 contraction at the level of the statistic: the hoisted
formulation the CUDA path uses (Sigma = L L^T, G = L^-1 T^T N^-1, w = C^-1 r), with Y = G [s c] computed the way the
INT8 tensor-core plan would -- signed 7-bit digit planes of G (per-row power-of-two scale), unsigned bit-field digits
of s/2 + 1/2, exact integer products per digit weight, offset removal, fp64 recombination -- and everything else in
fp64. It must meet the same envelope against the longdouble truth as the CUDA kernels do on the golden fixture,
including the ill-conditioned bins next to the red-noise Fourier frequencies. (tools/probes/ holds the GPU side.)"""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

from conftest import term_tolerance
from oracle import fp_oracle as o

NS, BITS = 8, 7


def _signed_digits(X, e):
    r = (X / np.exp2(e)).astype(np.longdouble)  # |r| <= 1/2
    out = []
    for i in range(1, NS + 1):
        w = np.longdouble(2.0) ** (BITS * i)
        d = np.rint(r * w)
        out.append(d.astype(np.int64))
        r = r - d / w
    return out


def _unsigned_digits(X):
    """base-128 digits of (x/2 + 1/2) 2^56 for x in [-1, 1]; the first digit may be 128"""
    q = np.floor((X.astype(np.longdouble) * 0.5 + 0.5) * np.longdouble(2.0) ** 56 + 0.5)
    out = []
    for _ in range(NS - 1):
        hi = np.floor(q / 128)
        out.append((q - hi * 128).astype(np.int64))
        q = hi
    out.append(q.astype(np.int64))
    return out[::-1]


def _split_product(G, S):
    """Y = G S (m x n times n x k) through digit planes; integer arithmetic is exact, as int32 accumulation is."""
    e = np.ceil(np.log2(np.abs(G).max(axis=1, keepdims=True))) + 1
    dG, dS = _signed_digits(G, e), _unsigned_digits(S)
    assert max(np.abs(d).max() for d in dG) <= 64 and max(d.max() for d in dS) <= 128 and min(d.min() for d in dS) >= 0
    acc = [np.zeros((G.shape[0], S.shape[1]), dtype=np.int64) for _ in range(NS)]
    for i in range(NS):
        for j in range(NS - i):
            acc[i + j] += dG[i] @ dS[j]
    assert max(np.abs(a).max() for a in acc) < 2 ** 31  # fits the int32 accumulators of the MMA
    y = np.zeros(acc[0].shape)
    for g in range(NS - 1, -1, -1):  # smallest weight first, fp64 as in the epilogue
        y = y + acc[g].astype(np.float64) * 2.0 ** (-BITS * (g + 2))
    # offset: every S entry carried +1/2, i.e. (1/2) sum_k g_jk per row, exact from the digits
    gsum = sum(dG[i].sum(axis=1).astype(np.float64) * 2.0 ** (-BITS * (i + 1)) for i in range(NS - 1, -1, -1))
    return (y - 0.5 * gsum[:, None]) * np.exp2(e + 1)


def _terms(g, product):
    freqs = g["freqs"]
    out = np.zeros((len(g.psrs), len(freqs)))
    for p, (q, Nvec, T, sigma) in enumerate(zip(g.psrs, g.lst("Nvec"), g.lst("T"), g.lst("sigma"))):
        t, r = q.toas, q.residuals
        L = np.linalg.cholesky(sigma)
        G = solve_triangular(L, (T / Nvec[:, None]).T, lower=True)
        w = r / Nvec - G.T @ (G @ r)
        ph = (2 * np.pi * freqs)[None, :] * t[:, None]  # ((2 pi) f) t, reference rounding order
        s, c = np.sin(ph), np.cos(ph)
        Y = product(G, np.concatenate((s, c), axis=1))
        Ys, Yc = Y[:, : len(freqs)], Y[:, len(freqs):]
        ninv = (1.0 / Nvec)[:, None]
        m00 = (s * s * ninv).sum(0) - (Ys * Ys).sum(0)
        m01 = (s * c * ninv).sum(0) - (Ys * Yc).sum(0)
        m11 = (c * c * ninv).sum(0) - (Yc * Yc).sum(0)
        n0, n1 = (s * w[:, None]).sum(0), (c * w[:, None]).sum(0)
        det = m00 * m11 - m01 * m01
        out[p] = 0.5 * (n0 * (m11 * n0 - m01 * n1) + n1 * (m00 * n1 - m01 * n0)) / det
    return out


@pytest.mark.parametrize("name", ["fp_white", "fp_red"])
def test_split_precision_contraction_meets_the_parity_envelope(golden, name):
    g = golden(name)
    ora = o.fp_sweep(g["freqs"], g.lst("toas"), g.lst("res"), g.lst("Nvec"), g.lst("T"), g.lst("sigma"), per_pulsar=True)
    tol = term_tolerance(g["truth_terms"], g["cond"], ora)
    plain = _terms(g, lambda G, S: G @ S)
    split = _terms(g, _split_product)
    # the emulation harness itself (hoisted formulation in NumPy, plain fp64 product) is inside the envelope ...
    assert np.all(np.abs(plain - g["truth_terms"]) <= 2 * tol)
    # ... and so is the digit-plane product; it is not further from the truth than the plain fp64 product
    assert np.all(np.abs(split - g["truth_terms"]) <= 2 * tol)
    e_split = np.abs(split - g["truth_terms"]) / tol
    e_plain = np.abs(plain - g["truth_terms"]) / tol
    assert np.median(e_split) <= 1.5 * np.median(e_plain) + 1e-3
