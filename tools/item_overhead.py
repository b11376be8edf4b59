"""Fixed cost per work item (pulsar, 64-frequency tile) of the sweep kernel: same pulsar count and grid
(P=8, 9472 frequencies = 148 tiles per pulsar -> exactly 8 items per CTA), TOA count varied, so
time = 8 * (nch * t_chunk + t_item). Prints a least-squares fit."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth
P, F = 8, 9472
fr = torch.tensor(synth.fp_freqs(F), dtype=torch.float64, device="cuda")
rows = []
for n in (1280, 2560, 5120, 10240):
    pta = synth.make_pta(P, n)
    fp = fastfp_b200.FastFp(pta.psrs)
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    t0 = time.time()
    while time.time() - t0 < 0.8:
        fp(fr, *a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fp(fr, *a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rows.append((n // 32, ms))
    print(f"n={n} nch={n//32}: {ms:.3f} ms per sweep, {ms/8*1e3:.1f} us per item", flush=True)
x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows]) / 8 * 1e3
A = np.stack([x, np.ones_like(x)], 1)
(tc, ti), *_ = np.linalg.lstsq(A, y, rcond=None)
print(f"t_chunk = {tc:.3f} us ({tc*1965:.0f} cycles), t_item = {ti:.1f} us ({ti/tc:.1f} chunk times)")
