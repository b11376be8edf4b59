"""Stage timeline of the tensor-core sweep (diagnostic build only: FASTFP_B200_NVCC_FLAGS=-DFFP_I8_TRACE).

Runs one C2-shaped sweep with FASTFP_B200_I8_TRACE=<file> and prints, for CTA 0's first stages, the SM clock of
  0 MMAs of the stage complete (seen by the TMA thread)   1 issuer sees the stage complete   2 issuer has issued
  3 fp64 part starts   4 fp64 part done   5 planes stored
relative to the stage's own 'planes stored' time, plus the steady-state period per stage.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = os.environ.setdefault("FASTFP_B200_I8_TRACE", "gpurun_out/i8_trace.txt")
import torch
import fastfp_b200
from fastfp_b200 import synth

pta = synth.make_config("C2")
a = (pta.Nvecs, pta.Ts, pta.sigmas)
fd = torch.from_numpy(synth.fp_freqs(10_000)).cuda()
fp = fastfp_b200.FastFp(pta.psrs, path="i8")
fp(fd, *a); torch.cuda.synchronize()
t = np.loadtxt(out)
names = ["mma_done", "iss_sees", "iss_done", "A_start", "A_done", "B_done"]
k0, k1 = 40, 140
base = t[k0, 5]
print("stage " + " ".join(f"{n:>9s}" for n in names))
for k in range(k0, k0 + 24):
    print(f"{k:5d} " + " ".join(f"{t[k, e] - base:9.0f}" for e in range(6)))
per = (t[k1, 5] - t[k0, 5]) / (k1 - k0)
print(f"period per stage {per:.0f} cycles")
d = lambda e1, e0, sh=0: np.mean(t[k0 + sh:k1 + sh, e1] - t[k0:k1, e0])
print(f"A (fp64 part) duration           {d(4, 3):8.0f}")
print(f"B (integer part) duration        {d(5, 4):8.0f}")
print(f"A_done(k+1) -> issuer sees k     {np.mean(t[k0:k1, 1] - t[k0 + 1:k1 + 1, 4]):8.0f}")
print(f"issuer sees -> issued            {d(2, 1):8.0f}")
print(f"issuer sees -> MMAs complete     {d(0, 1):8.0f}")
print(f"MMAs complete(k) -> A_start(k+2) {np.mean(t[k0 + 2:k1 + 2, 3] - t[k0:k1, 0]):8.0f}")
print(f"MMAs complete(k) -> A_done(k+2)  {np.mean(t[k0 + 2:k1 + 2, 4] - t[k0:k1, 0]):8.0f}")
