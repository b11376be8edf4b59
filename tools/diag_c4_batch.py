"""Diagnostic: is a bin of the C4 sweep independent of its batch at F = 1e6? (bench.py check failed at 1 GPU)"""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastfp_b200
from fastfp_b200 import synth

pta = synth.make_config("C4")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
grid = synth.fp_freqs(F)
dev = torch.device("cuda", 0)
fd = torch.from_numpy(grid).to(dev)
fp = fastfp_b200.FastFp(pta.psrs)
a = (pta.Nvecs, pta.Ts, pta.sigmas)
full1 = fp(fd, *a); torch.cuda.synchronize()
full2 = fp(fd, *a); torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(full1).all()), "run-to-run equal:", bool(torch.equal(full1, full2)),
      "n diff", int((full1 != full2).sum()))
for lo in (0, 4096, F // 2, F // 2 + 37, F - 4096):
    sub = fp(fd[lo:lo + 4096].contiguous(), *a)
    d = (sub != full1[lo:lo + 4096])
    nd = int(d.sum())
    print("slice", lo, "n diff", nd, end=" ")
    if nd:
        idx = torch.nonzero(d).flatten()[:8].cpu().numpy()
        rel = ((sub - full1[lo:lo + 4096]).abs() / sub.abs()).max().item()
        print("first idx", idx, "max rel", rel)
    else:
        print()
# per-pulsar terms of the differing bins
lo = F // 2
t_full = torch.empty((pta.P, F), dtype=torch.float64, device=dev)
pack = fp.prepare(*a)
pack.fp_sweep((fd.data_ptr(), F), out=t_full.data_ptr(), terms=True); torch.cuda.synchronize()
t_sub = torch.empty((pta.P, 4096), dtype=torch.float64, device=dev)
s = fd[lo:lo + 4096].contiguous()
pack.fp_sweep((s.data_ptr(), 4096), out=t_sub.data_ptr(), terms=True); torch.cuda.synchronize()
dt = (t_sub != t_full[:, lo:lo + 4096])
print("terms differing:", int(dt.sum()), "pulsars with diffs:", torch.nonzero(dt.any(1)).flatten().cpu().numpy()[:20])
# is the sum order the issue? compare ordered sums
acc = torch.zeros(4096, dtype=torch.float64, device=dev)
for p in range(pta.P):
    acc = acc + t_sub[p]
print("ordered sum of sub terms == sub sweep:", bool(torch.equal(acc, fp(s, *a))))
acc2 = torch.zeros(4096, dtype=torch.float64, device=dev)
for p in range(pta.P):
    acc2 = acc2 + t_full[p, lo:lo + 4096]
print("ordered sum of full terms == full sweep slice:", bool(torch.equal(acc2, full1[lo:lo + 4096])))
