"""Per-call latency of the reference's own small case (BASELINE configs[0]: 10 pulsars x 1000 TOAs,
white noise only, ONE frequency, as examples/run_fp.py is called for a single fgw) through the public API
with host buffers, plus the 100-frequency call. usage: latency_small.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth
pta = synth.make_config("C1")
fp = fastfp_b200.FastFp(pta.psrs)
mats = (pta.Nvecs, pta.Ts, pta.sigmas)
for F in (1, 100, 10000):
    fr = synth.fp_freqs(F) if F > 1 else np.float64(2e-8)
    for _ in range(20): fp(fr, *mats)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 200
    for _ in range(reps): out = fp(fr, *mats)
    dt = (time.perf_counter() - t0) / reps
    print(f"C1 F={F}: {dt*1e6:.1f} us per call (host in, host out), {F*pta.P/dt:.4g} evals/s")
