"""Diagnostic: locate the non-finite bins of the 1e6-frequency C4 sweep."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastfp_b200
from fastfp_b200 import synth
from oracle import fp_oracle as o

pta = synth.make_config("C4")
F = 1_000_000
grid = synth.fp_freqs(F)
dev = torch.device("cuda", 0)
fd = torch.from_numpy(grid).to(dev)
fp = fastfp_b200.FastFp(pta.psrs)
a = (pta.Nvecs, pta.Ts, pta.sigmas)
pack = fp.prepare(*a)
for rep in range(3):
    t = torch.empty((pta.P, F), dtype=torch.float64, device=dev)
    pack.fp_sweep((fd.data_ptr(), F), out=t.data_ptr(), terms=True); torch.cuda.synchronize()
    bad = torch.nonzero(~torch.isfinite(t))
    print("rep", rep, "non-finite terms:", bad.shape[0], bad[:10].cpu().numpy().tolist())
    for p, f in bad[:6].cpu().numpy().tolist():
        print("  pulsar", p, "bin", f, "freq %.17g" % grid[f], "value", t[p, f].item(), "tile", f // 64, "pos", f % 64,
              "neighbours", t[p, f - 1].item(), t[p, f + 1].item())
        sub = torch.empty((pta.P, 129), dtype=torch.float64, device=dev)
        s = fd[f - 64:f + 65].contiguous()
        pack.fp_sweep((s.data_ptr(), 129), out=sub.data_ptr(), terms=True); torch.cuda.synchronize()
        print("   recomputed in a 129-bin batch:", sub[p, 64].item(),
              "oracle:", o.fp_sweep(grid[f:f + 1], [pta.toas[p]], [pta.residuals[p]], [pta.Nvecs[p]], [pta.Ts[p]], [pta.sigmas[p]])[0])
