import os, sys, subprocess
code = '''
import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import fastfp_b200
from fastfp_b200 import synth
pta = synth.make_pta(8, 5000)
fp = fastfp_b200.FastFp(pta.psrs)
fr = torch.tensor(synth.fp_freqs(9472), dtype=torch.float64, device="cuda")
for _ in range(40): fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("FASTFP_DBG=%s: %.3f ms  cycles/iter=%.0f" % (os.environ.get("FASTFP_DBG","0"), ms, ms*1e-3*1.965e9/(8*157)))
'''
for d in ("0", "1", "2"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTFP_DBG=d))
