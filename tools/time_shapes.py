import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth
def run(P, n, F, reps=3):
    pta = synth.make_pta(P, n)
    fp = fastfp_b200.FastFp(pta.psrs)
    fr = torch.tensor(synth.fp_freqs(F), dtype=torch.float64, device="cuda")
    fp(fr, pta.Nvecs, pta.Ts, pta.sigmas); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nch = (n + 31) // 32
    waves = P * ((F + 63) // 64) / 148
    print(f"P={P} n={n} F={F}: {ms:.3f} ms, waves={waves:.2f}, cycles/iter@1.93GHz={ms*1e-3*1.93e9/(np.ceil(waves)*nch):.0f}", flush=True)
for args in [(2, 5000, 9472), (8, 5000, 9472), (45, 5000, 9472), (45, 5000, 10000), (8, 5000, 2048), (8, 5000, 64*37)]:
    run(*args)
