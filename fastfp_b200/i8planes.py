"""Host-side digit planes for the INT8 (split-precision) form of the sweep contraction -- scaffolding for the
round-2 kernel (tools/probes/README.md). Not used by the product path.

``Y = G [s c]`` is evaluated as 36 INT8 tensor-core products of 7-bit digit planes with exact int32 accumulation:

* ``G`` (m x n, one pulsar): per row a power-of-two scale ``2^e`` with ``|G / 2^e| <= 1/2`` and 8 signed digits
  ``d_i in [-64, 64]``, ``G / 2^e = sum_i d_i 2^(-7 i)`` (``i = 1..8``), the remainder below ``2^-57``;
* ``s, c in [-1, 1]``: 8 unsigned base-128 digits of ``(x/2 + 1/2) 2^56`` (the first may be 128); the ``+1/2`` adds
  ``(1/2) sum_i g_ji`` to row ``j`` of the product, which :func:`g_digit_planes` returns from the integer digit sums.

The planes are laid out the way ``tcgen05.mma`` reads a K-major operand from shared memory (validated on a B200 by
``tools/probes/umma_i8_split_check.cu``): stages of ``KB`` TOAs (= bytes per row), rows in groups of 8, the
16-byte chunk index XORed with the low row bits (SWIZZLE_128B / 64B / 32B for ``KB`` = 128 / 64 / 32).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

NS, BITS, ROWS = 8, 7, 128


def swz_offset(r, c, KB: int):
    """Byte offset of (row ``r``, K byte ``c``) inside one swizzled K-major tile with ``KB``-byte rows."""
    if KB not in (128, 64, 32):
        raise ValueError("KB must be 128, 64 or 32")
    r, c = np.asarray(r), np.asarray(c)
    shift = {128: 0, 64: 1, 32: 2}[KB]
    return (r >> 3) * (8 * KB) + (r & 7) * KB + ((((c >> 4) ^ ((r & 7) >> shift))) << 4) + (c & 15)


def signed_digits(X, e) -> np.ndarray:
    """``X / 2^e`` (``|.| <= 1/2``) -> ``(NS, *X.shape)`` int8 digits, most significant first."""
    r = np.ldexp(np.asarray(X, dtype=np.longdouble), -np.asarray(e, dtype=np.int64))
    out = np.empty((NS,) + r.shape, dtype=np.int8)
    for i in range(1, NS + 1):
        w = np.longdouble(2.0) ** (BITS * i)
        d = np.rint(r * w)
        out[i - 1] = d.astype(np.int8)
        r = r - d / w
    return out


def unsigned_digits(x) -> np.ndarray:
    """``x in [-1, 1]`` -> ``(NS, *x.shape)`` uint8 base-128 digits of ``(x/2 + 1/2) 2^56``, most significant first."""
    q = np.floor((np.asarray(x, dtype=np.longdouble) * 0.5 + 0.5) * np.longdouble(2.0) ** 56 + 0.5)
    out = np.empty((NS,) + q.shape, dtype=np.uint8)
    for i in range(NS - 1, 0, -1):
        hi = np.floor(q / 128)
        out[i] = (q - hi * 128).astype(np.uint8)
        q = hi
    out[0] = q.astype(np.uint8)
    return out


def g_digit_planes(G, KB: int = 32) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Digit planes of one pulsar's ``G`` (m x n, m <= 128).

    Returns ``(planes, e, scale, roff)``: ``planes`` int8 ``(nstage, NS, ROWS * KB)`` in the swizzled operand layout
    (TOAs padded with zeros to a multiple of ``KB``, rows ``>= m`` zero); ``e`` the row exponents; ``scale`` =
    ``2^(e+1)`` (the factor that undoes ``G / 2^e`` and ``x / 2``); ``roff`` = ``(1/2) sum_i g_ji`` in units of
    ``2^e``, exact from the integer digit sums, to subtract from the recombined accumulators before scaling."""
    G = np.asarray(G, dtype=np.float64)
    m, n = G.shape
    if m > ROWS:
        raise ValueError(f"at most {ROWS} basis rows per MMA operand")
    mx = np.abs(G).max(axis=1)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0).astype(np.int64)
    # ceil(log2) can be one short when max|G| is an exact power of two times (1 + tiny): enforce |G / 2^e| <= 1/2
    e = np.where(np.ldexp(mx, -e) > 0.5, e + 1, e)
    d = signed_digits(G, e[:, None])                       # (NS, m, n)
    nstage = -(-n // KB)
    full = np.zeros((NS, ROWS, nstage * KB), dtype=np.int8)
    full[:, :m, :n] = d
    rr, cc = np.meshgrid(np.arange(ROWS), np.arange(KB), indexing="ij")
    off = swz_offset(rr, cc, KB).ravel()
    planes = np.zeros((nstage, NS, ROWS * KB), dtype=np.int8)
    tiles = full.reshape(NS, ROWS, nstage, KB).transpose(2, 0, 1, 3).reshape(nstage, NS, ROWS * KB)
    planes[:, :, off] = tiles
    sums = d.astype(np.int64).sum(axis=2)                  # (NS, m), exact
    gsum = np.zeros(ROWS)
    for i in range(NS - 1, -1, -1):
        gsum[:m] += sums[i].astype(np.float64) * 2.0 ** (-BITS * (i + 1))
    e_full = np.zeros(ROWS, dtype=np.int64)
    e_full[:m] = e
    return planes, e_full, np.exp2(e_full + 1.0), 0.5 * gsum
