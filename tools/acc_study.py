"""Accuracy of the sweep vs longdouble truth at ill-conditioned points, with and without the
blocked (two-level) accumulation (FASTFP_DBG=4 disables the level-1 flush)."""
import os, sys, subprocess, json
code = '''
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import fastfp_b200
from fastfp_b200 import synth
from oracle import fp_oracle, truth
out = {}
for tag, P, n in (("n5000", 6, 5000), ("n10000", 6, 10000)):
    pta = synth.make_pta(P, n, seed=synth.SEED0 + 12)   # pulsars 12..17 include the steep-red-noise ones
    k = np.arange(1, 9)
    fr = np.concatenate((synth.fp_freqs(10000)[[0, 1, 2, 5, 17, 40, 100]], k / pta.Tspan, (k + 0.37) / pta.Tspan))
    fp = fastfp_b200.FastFp(pta.psrs)
    got = fp.per_pulsar_terms(fr, pta.Nvecs, pta.Ts, pta.sigmas)
    tt, cond = truth.fp_sweep_truth(fr, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, pta.sigmas)
    ora = fp_oracle.fp_sweep(fr, pta.toas, pta.residuals, pta.Nvecs, pta.Ts, pta.sigmas, per_pulsar=True)
    tt = tt.astype(np.float64)
    eg = np.abs(got - tt) / (2.2e-16 * cond); eo = np.abs(ora - tt) / (2.2e-16 * cond)
    out[tag] = dict(gpu_max=float(eg.max()), gpu_rms=float(np.sqrt((eg**2).mean())), ora_max=float(eo.max()), ora_rms=float(np.sqrt((eo**2).mean())),
                    kappa_max=float((cond/np.abs(tt)).max()), gpu_rel_max=float(np.abs(got/tt-1).max()), ora_rel_max=float(np.abs(ora/tt-1).max()))
print(json.dumps(out))
'''
for d in ("0", "4"):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTFP_DBG=d), capture_output=True, text=True)
    print("FASTFP_DBG=" + d, r.stdout.strip()[-900:], r.stderr[-300:])
