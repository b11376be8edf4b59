"""Timing of the noise-marginalised path. usage: time_nmfp.py CONFIG F D"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastfp_b200 import NMFP, CURN_container, RN_container, synth
cfg, F, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pta = synth.make_config(cfg)
curn = CURN_container(pta.Ffreqs)
sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
nm = NMFP(pta.psrs, sigs)
samples = synth.draw_samples(pta, D)
fr = torch.tensor(synth.nmfp_freqs(F, pta.Tspan), dtype=torch.float64, device="cuda")
t0 = time.time(); nm.prepare(pta.Nvecs, pta.Ts, pta.TNTs); torch.cuda.synchronize(); print("pack s", time.time() - t0)
t1 = time.time()
while time.time() - t1 < 1.5:  # clock spin-up: the SM clock needs ~0.4 s of load to leave idle
    out = nm(fr, samples, pta.Nvecs, pta.Ts, pta.TNTs); torch.cuda.synchronize()
reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): out = nm(fr, samples, pta.Nvecs, pta.Ts, pta.TNTs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{cfg} nmfp F={F} D={D}: {ms:.2f} ms per sweep, {F*D*pta.P/ms*1e3:.4g} evals/s (freq x psr x draw)", out.shape)
