"""Bring-up check of the INT8 tensor-core sweep: i8 vs fp64 DMMA vs oracle/truth on small cases, then timing on C2."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import fastfp_b200
from fastfp_b200 import synth
from oracle import fp_oracle as o, truth
from conftest import term_tolerance

def case(pta, freqs, label):
    a = (pta.Nvecs, pta.Ts, pta.sigmas)
    f8 = fastfp_b200.FastFp(pta.psrs, path="i8")
    f64 = fastfp_b200.FastFp(pta.psrs, path="fp64")
    t8 = f8.per_pulsar_terms(freqs, *a)
    t64 = f64.per_pulsar_terms(freqs, *a)
    args = (freqs, pta.toas, pta.residuals, *a)
    ora = o.fp_sweep(*args, per_pulsar=True)
    tt, cond = truth.fp_sweep_truth(*args)
    tol = term_tolerance(tt.astype(float), cond, ora)
    e8 = np.abs(t8 - tt.astype(float)) / tol
    e64 = np.abs(t64 - tt.astype(float)) / tol
    print(f"{label}: path {f8.prepare(*a).path}/{f64.prepare(*a).path}  i8 err/tol max {e8.max():.3g}  fp64 err/tol max {e64.max():.3g}  "
          f"rel i8 vs fp64 max {np.nanmax(np.abs(t8 / t64 - 1)):.3g}  finite {np.isfinite(t8).all()}")
    if not (e8.max() < 1):
        bad = np.argwhere(~(e8 < 1))[:10]
        for p, f in bad:
            print("   bad", p, f, t8[p, f], t64[p, f], float(tt[p, f]))
    return e8.max() < 1

ok = True
pta = synth.make_pta(3, [200, 333, 257], n_tm=[8, 12, 10], ncomps=30, seed=4242)
ok &= case(pta, synth.fp_freqs(33), "small ragged (F=33)")
ok &= case(pta, synth.fp_freqs(1), "F=1")
pta = synth.make_pta(2, [64, 1000], n_tm=[2, 5], ncomps=3, seed=7)
ok &= case(pta, np.concatenate((synth.fp_freqs(70), np.array([1.0, 2.5, 7.0]) / pta.Tspan)), "m=8/11, F=73")
pta = synth.make_pta(3, [2500, 1801, 3000], n_tm=[40, 12, 60], ncomps=30, seed=77)
ok &= case(pta, synth.fp_freqs(100), "m=100/72/120, n~2500")
print("ALL OK" if ok else "FAILURES")
# timing on C2
pta = synth.make_config("C2")
a = (pta.Nvecs, pta.Ts, pta.sigmas)
fd = torch.from_numpy(synth.fp_freqs(10_000)).cuda()
for path in ("i8", "fp64"):
    fp = fastfp_b200.FastFp(pta.psrs, path=path)
    out = fp(fd, *a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = fp(fd, *a)
    e1.record(); torch.cuda.synchronize()
    print(f"C2 {path}: {e0.elapsed_time(e1) / 5:.3f} ms per sweep; sum {out.sum().item():.10g}")
    if path == "i8":
        r8 = out.clone()
    else:
        print("C2 i8 vs fp64 max rel", ((r8 - out).abs() / out.abs()).max().item())
