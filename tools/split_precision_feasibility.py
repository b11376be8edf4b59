"""Feasibility study for round 2: the sweep's contraction Y = G [s c] through an
error-free split into 7-bit slices -- what INT8 tensor-core MMAs with exact int32 accumulation would compute --
emulated with integers on the CPU. Reports, for the C2-like pulsar, the error of the split product for s slices
against a long-double product, next to the error of the plain fp64 product, both in units of
eps * sum_k |G_jk| |S_kf| (the natural scale of a dot product's rounding error).
usage: split_precision_feasibility.py [n] [F]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastfp_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
F = int(sys.argv[2]) if len(sys.argv) > 2 else 24
pta = synth.make_pta(1, n)
T, Nvec, sigma, t = pta.Ts[0], pta.Nvecs[0], pta.sigmas[0], pta.toas[0]
L = np.linalg.cholesky(sigma)
G = np.linalg.solve(L, (T / Nvec[:, None]).T)            # m x n, what the packets hold
freqs = synth.fp_freqs(10_000)[:: 10_000 // F][:F]
ph = (2 * np.pi * freqs)[None, :] * t[:, None]
S = np.concatenate((np.sin(ph), np.cos(ph)), axis=1)      # n x 2F
Gl, Sl = G.astype(np.longdouble), S.astype(np.longdouble)
Y_true = Gl @ Sl
scale = (np.abs(Gl) @ np.abs(Sl)).astype(float)          # sum |G||S| per output
eps = np.finfo(float).eps
err64 = np.abs((G @ S) - Y_true).astype(float) / scale / eps
print(f"n={n} m={G.shape[0]} F={F}: plain fp64 product  max err {err64.max():.3f}  median {np.median(err64):.3f}  [eps * sum|G||S|]")

BITS = 7
def digits(X, e, s):
    """X / 2^e (|.| <= 1, e per row or scalar) as s signed 7-bit digits d_i with X/2^e ~ sum d_i 2^(-7 i)"""
    r = (X / np.exp2(e)).astype(np.longdouble)
    out = []
    for i in range(1, s + 1):
        d = np.rint(r * np.longdouble(2.0) ** (BITS * i))
        out.append(d.astype(np.int64))
        r = r - d * np.longdouble(2.0) ** (-BITS * i)
    return out
eG = np.ceil(np.log2(np.abs(G).max(axis=1, keepdims=True))) + 1   # per-row scale of G: |G / 2^e| <= 1/2
for s in (5, 6, 7, 8, 9):
    dG, dS = digits(G, eG, s), digits(S, 1.0, s)                  # |S / 2| <= 1/2
    assert max(np.abs(d).max() for d in dG + dS) <= 64
    acc = np.zeros(Y_true.shape, dtype=np.longdouble)
    nprod = 0
    for i in range(s):
        for j in range(s - i):                                # slices with i + j < s: the rest is below the target
            acc += (dG[i] @ dS[j]).astype(np.longdouble) * np.longdouble(2.0) ** (-BITS * (i + j + 2))
            nprod += 1
    Y = acc * np.exp2(eG + 1.0).astype(np.longdouble)
    e = np.abs(Y - Y_true).astype(float) / scale / eps
    print(f"  {s} slices, {nprod:2d} int8 products: max err {e.max():10.3f}  median {np.median(e):8.3f}")
print("int32 accumulation is exact for K <= 2^(31-14) = 131072 terms per MMA chain; products needed ~ s(s+1)/2")
