import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastfp_b200 import NMFP, CURN_container, RN_container, synth
pta = synth.make_config("C3")
curn = CURN_container(pta.Ffreqs)
sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
nm = NMFP(pta.psrs, sigs)
samples = synth.draw_samples(pta, 1000)
fr = torch.tensor(synth.nmfp_freqs(1000, pta.Tspan), dtype=torch.float64, device="cuda")
for _ in range(3): nm(fr, samples, pta.Nvecs, pta.Ts, pta.TNTs)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    nm(fr, samples, pta.Nvecs, pta.Ts, pta.TNTs); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
