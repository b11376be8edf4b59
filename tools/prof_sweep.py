"""One plain-Fp sweep at a chosen size, for ncu / quick timing. usage: prof_sweep.py CONFIG F [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastfp_b200
from fastfp_b200 import synth
cfg, F = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pta = synth.make_config(cfg)
fp = fastfp_b200.FastFp(pta.psrs)
fr = torch.tensor(synth.fp_freqs(F), dtype=torch.float64, device="cuda")
res = fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): res = fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = sum(2.0 * (2 * m + 5) * n for n, m in zip(fp._pack.n, fp._pack.m)) * F
print(f"{cfg} F={F}: {ms:.3f} ms  {F*pta.P/ms*1e3:.4g} evals/s  {flops/ms/1e9:.2f} TFLOP/s(gemm+scalars)")
