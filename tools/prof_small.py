import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fastfp_b200
from fastfp_b200 import synth
pta = synth.make_pta(4, 5000)
fp = fastfp_b200.FastFp(pta.psrs)
fr = torch.tensor(synth.fp_freqs(9472), dtype=torch.float64, device="cuda")
for _ in range(2): fp(fr, pta.Nvecs, pta.Ts, pta.sigmas)
torch.cuda.synchronize()
