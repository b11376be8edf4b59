"""CPU emulation of the tensor-core contraction of csrc/fp_sweep_i8.cu at the level of the statistic.

The hoisted formulation the CUDA paths use (Sigma = L L^T, G = L^-1 T^T N^-1, w = C^-1 r) with Y = G [s c] computed
exactly the way the INT8 kernel does it -- radix-256 balanced digit planes of G (per-row power-of-two scale, |g| < 1/4,
Qg = rint(g 2^55)) and of sin/cos (Q = rint(x 2^54), as ``digits7`` extracts them), planes stored in the SWIZZLE_32B K-major operand layout and read back through it, 28 exact integer plane products
into 7 accumulators, Horner recombination in fp64, the w row supplying (s|r), (c|r) -- and everything else in fp64. It
must meet the same envelope against the longdouble truth as the CUDA kernels do on the golden fixtures, including the
ill-conditioned bins next to the red-noise Fourier frequencies; and the digit formats' invariants are checked directly.
"""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

from conftest import term_tolerance
from oracle import fp_oracle as o

NPL, KT = 7, 32
BIAS = 0x0080808080808080


def swz32(r, c):
    """byte offset of (row, K byte) in a K-major tile with 32-byte rows, SWIZZLE_32B (fp_sweep_i8.cu::swz32)"""
    r, c = np.asarray(r), np.asarray(c)
    return (r >> 3) * 256 + (r & 7) * 32 + ((((c >> 4) ^ ((r & 7) >> 2)) & 1) << 4) + (c & 15)


def digits7(x):
    """fp_sweep_i8.cu::digits7 in NumPy: x in [-1, 1] -> (7, ...) int8 balanced digits of Q = rint(x 2^54), most
    significant first (the kernel scales by an exponent add and rounds with one F2I.S64.F64)."""
    Q = np.rint(np.ldexp(np.asarray(x, dtype=np.float64), 54)).astype(np.int64)
    U = Q + BIAS
    return np.stack([(((U >> (8 * (6 - p))) & 0xff) ^ 0x80).astype(np.uint8).view(np.int8) for p in range(NPL)])


def g_planes(G):
    """i8_rowscale_kernel + i8_planes_kernel: row exponents (|G / 2^e| < 1/4), digits of rint(G 2^(55 - e))"""
    mx = np.abs(G).max(axis=1)
    e = np.where(mx > 0, np.frexp(np.where(mx > 0, mx, 1.0))[1] + 2, 0)
    Q = np.rint(np.ldexp(G, (55 - e)[:, None])).astype(np.int64)
    assert np.abs(Q).max() <= 2 ** 53
    U = (Q + BIAS) ^ BIAS
    return np.stack([((U >> (8 * (6 - p))) & 0xff).astype(np.uint8).view(np.int8) for p in range(NPL)]), e


def test_digit_formats_are_exact_representations():
    rng = np.random.default_rng(0)
    x = np.concatenate((rng.uniform(-1, 1, 20000), [0.0, 1.0, -1.0, 2.0 ** -60, -2.0 ** -30, 1 - 2.0 ** -53]))
    d = digits7(x).astype(np.int64)
    Q = sum(d[p] * 256 ** (6 - p) for p in range(NPL))
    exact = [np.longdouble(v) * np.longdouble(2.0) ** 54 for v in x]
    want = np.array([int(np.rint(v)) for v in exact], dtype=np.int64)
    np.testing.assert_array_equal(Q, want)  # the seven digits ARE x 2^54 rounded to nearest
    assert np.abs(x - Q * 2.0 ** -54).max() <= 2.0 ** -55  # i.e. sin/cos are quantised to within 2^-55
    G = rng.standard_normal((9, 300)) * 10.0 ** rng.uniform(-3, 3, (9, 1)) * 10.0 ** rng.uniform(-1, 1, (9, 300))
    planes, e = g_planes(G)
    Qg = sum(planes[p].astype(np.int64) * 256 ** (6 - p) for p in range(NPL))
    np.testing.assert_array_equal(Qg, np.rint(np.ldexp(G, (55 - e)[:, None])).astype(np.int64))
    assert np.abs(np.ldexp(G, -e[:, None])).max() < 0.25 and np.abs(planes[0]).max() <= 33


def _tensor_product(G, S):
    """Y = G S^T-like product (G: m x n, S: n x k) through the kernel's data path, stage by stage."""
    m, n = G.shape
    k = S.shape[1]
    rows = (m + 7) // 8 * 8
    gp, e = g_planes(G)
    nst = -(-n // KT)
    acc = [np.zeros((m, k), dtype=np.int64) for _ in range(NPL)]
    rr, cc = np.meshgrid(np.arange(rows), np.arange(KT), indexing="ij")
    offA = swz32(rr, cc)
    r2, c2 = np.meshgrid(np.arange(k), np.arange(KT), indexing="ij")
    offB = swz32(r2, c2)
    for st in range(nst):
        lo, hi = st * KT, min(n, (st + 1) * KT)
        A = np.zeros((NPL, rows * KT), dtype=np.int8)   # what the TMA copy lands in shared memory
        B = np.zeros((NPL, k * KT), dtype=np.int8)      # what the producers store
        tile = np.zeros((NPL, rows, KT), dtype=np.int8)
        tile[:, :m, : hi - lo] = gp[:, :, lo:hi]
        A[:, offA.ravel()] = tile.reshape(NPL, -1)
        stile = np.zeros((NPL, k, KT), dtype=np.int8)
        stile[:, :, : hi - lo] = digits7(S[lo:hi].T)
        B[:, offB.ravel()] = stile.reshape(NPL, -1)
        a = A[:, offA].astype(np.int64)[:, :m]          # what the MMA reads back through the descriptor layout
        b = B[:, offB].astype(np.int64)
        for i in range(NPL):
            for j in range(NPL - i):
                acc[i + j] += a[i] @ b[j].T
    assert max(np.abs(x).max() for x in acc) < 2 ** 31  # fits the int32 accumulators in tensor memory
    y = acc[NPL - 1].astype(np.float64)
    for g in range(NPL - 2, -1, -1):                    # the epilogue's Horner recombination, smallest weight first
        y = y * 0.00390625 + acc[g].astype(np.float64)
    return y * np.exp2(e - 13.0)[:, None]


def _terms(g, tensor):
    freqs = g["freqs"]
    out = np.zeros((len(g.psrs), len(freqs)))
    for p, (q, Nvec, T, sigma) in enumerate(zip(g.psrs, g.lst("Nvec"), g.lst("T"), g.lst("sigma"))):
        t, r = q.toas, q.residuals
        L = np.linalg.cholesky(sigma)
        G = solve_triangular(L, (T / Nvec[:, None]).T, lower=True)
        w = r / Nvec - G.T @ (G @ r)
        ph = (2 * np.pi * freqs)[None, :] * t[:, None]  # ((2 pi) f) t, reference rounding order
        s, c = np.sin(ph), np.cos(ph)
        S = np.concatenate((s, c), axis=1)
        ninv = (1.0 / Nvec)[:, None]
        if tensor:  # rows of G plus the w row through the digit planes
            Y = _tensor_product(np.concatenate((G, w[None, :]), axis=0), S)
            n0, n1 = Y[-1, : len(freqs)], Y[-1, len(freqs):]
            Y = Y[:-1]
        else:
            Y = G @ S
            n0, n1 = (s * w[:, None]).sum(0), (c * w[:, None]).sum(0)
        Ys, Yc = Y[:, : len(freqs)], Y[:, len(freqs):]
        sNs = (s * s * ninv).sum(0)
        # the tensor kernel's producers form s N^-1 s and s N^-1 c only: c N^-1 c = sum 1/N - s N^-1 s
        cNc = float((1.0 / Nvec.astype(np.longdouble)).sum()) - sNs if tensor else (c * c * ninv).sum(0)
        m00 = sNs - (Ys * Ys).sum(0)
        m01 = (s * c * ninv).sum(0) - (Ys * Yc).sum(0)
        m11 = cNc - (Yc * Yc).sum(0)
        det = m00 * m11 - m01 * m01
        out[p] = 0.5 * (n0 * (m11 * n0 - m01 * n1) + n1 * (m00 * n1 - m01 * n0)) / det
    return out


@pytest.mark.parametrize("name", ["fp_white", "fp_red"])
def test_tensor_contraction_meets_the_parity_envelope(golden, name):
    g = golden(name)
    ora = o.fp_sweep(g["freqs"], g.lst("toas"), g.lst("res"), g.lst("Nvec"), g.lst("T"), g.lst("sigma"), per_pulsar=True)
    tol = term_tolerance(g["truth_terms"], g["cond"], ora)
    plain = _terms(g, tensor=False)
    split = _terms(g, tensor=True)
    # the emulation harness itself (hoisted formulation in NumPy, plain fp64 product) is inside the envelope ...
    assert np.all(np.abs(plain - g["truth_terms"]) <= 2 * tol)
    # ... and so is the digit-plane product; it is not further from the truth than the plain fp64 product
    assert np.all(np.abs(split - g["truth_terms"]) <= 2 * tol)
    e_split = np.abs(split - g["truth_terms"]) / tol
    e_plain = np.abs(plain - g["truth_terms"]) / tol
    assert e_split.max() <= max(1.5 * e_plain.max(), 0.25)
    assert np.median(e_split) <= 1.5 * np.median(e_plain) + 1e-3
