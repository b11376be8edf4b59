import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastfp_b200
from fastfp_b200 import synth, _cabi
pta = synth.make_pta(2, 5000)
fp = fastfp_b200.FastFp(pta.psrs)
pack = fp.prepare(pta.Nvecs, pta.Ts, pta.sigmas)
F = 64 * 148  # every SM busy: realistic contention
tr = _cabi.debug_trace(pack, synth.fp_freqs(F))
np.save("gpurun_out/trace.npy", tr)
t0 = tr[:, :, 0].min(axis=1, keepdims=True)
print("iteration length (cycles), chunks 8..15:", (tr[9:17, 0, 0] - tr[8:16, 0, 0]))
for c in (10, 11, 30):
    print(f"chunk {c}: per warp [phase1, phase2, barrier wait]")
    for w in range(8):
        a = tr[c, w]
        print(f"  warp {w} start+{a[0]-t0[c,0]:6d}  p1 {a[1]-a[0]:6d}  p2 {a[2]-a[1]:6d}  wait {a[3]-a[2]:6d}")
