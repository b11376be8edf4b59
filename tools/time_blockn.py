"""Cost of the block-diagonal-N (kernel ECORR) variant of the sweep next to the diagonal-N one, same
pulsars and grid. usage: time_blockn.py P N F [TOAS_PER_EPOCH]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import BlockNvec, synth
P, n, F = (int(v) for v in sys.argv[1:4])
per = int(sys.argv[4]) if len(sys.argv) > 4 else 4
pta = synth.make_pta(P, n)
rng = np.random.default_rng(5)
blocks, sig_block = [], []
for p in range(P):
    sl = [slice(a, a + per) for a in range(0, n - per + 1, per)]
    B = BlockNvec(pta.Nvecs[p], sl, rng.uniform(0.3, 3.0, len(sl)) * 1e-13)
    T = pta.Ts[p]
    TNT = T.T @ B.solve(T)
    blocks.append(B)
    sig_block.append(0.5 * (TNT + TNT.T) + np.diag(1.0 / pta.phis[p]))
fr = torch.tensor(synth.fp_freqs(F), dtype=torch.float64, device="cuda")
for label, Nv, sg in (("diagonal N", pta.Nvecs, pta.sigmas), (f"block N ({n // per} epochs of {per})", blocks, sig_block)):
    fp = fastfp_b200.FastFp(pta.psrs)
    t0 = time.time(); fp.prepare(Nv, pta.Ts, sg); torch.cuda.synchronize(); tp = time.time() - t0
    t1 = time.time()
    while time.time() - t1 < 1.0:
        fp(fr, Nv, pta.Ts, sg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = fp(fr, Nv, pta.Ts, sg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label}: pack {tp:.2f} s, {ms:.2f} ms per sweep, {F * P / ms * 1e3:.4g} evals/s, finite={bool(torch.isfinite(out).all())}")
