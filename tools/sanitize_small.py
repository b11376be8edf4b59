import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastfp_b200
from fastfp_b200 import NMFP, CURN_container, RN_container, synth
pta = synth.make_pta(2, [700, 333], n_tm=[12, 9], ncomps=30, seed=3)
f = synth.fp_freqs(70)
fp = fastfp_b200.FastFp(pta.psrs)
a = fp(f, pta.Nvecs, pta.Ts, pta.sigmas)
print("sweep path:", fp.prepare(pta.Nvecs, pta.Ts, pta.sigmas).path)   # auto: the tensor-core kernel for this pack
a64 = fastfp_b200.FastFp(pta.psrs, path="fp64")(f, pta.Nvecs, pta.Ts, pta.sigmas)
print("max rel dev between the two kernels:", float(np.nanmax(np.abs(a / a64 - 1))))
curn = CURN_container(pta.Ffreqs[:10])
sigs = [RN_container(q, Ffreqs=pta.Ffreqs, add_curn=True, curn_container=curn) for q in pta.psrs]
b = NMFP(pta.psrs, sigs)(f[:40], synth.draw_samples(pta, 9), pta.Nvecs, pta.Ts, pta.TNTs)
x = fastfp_b200.get_xCy(pta.Nvecs[0], pta.Ts[0], pta.sigmas[0], pta.psrs[0].toas * 1e-9, pta.psrs[0].residuals)
print("ok", a[:2], b.shape, x)
