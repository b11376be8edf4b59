"""fp64-pipe probes (fastfp_fp64_peak kinds): 1 DMMA peak, 0 DFMA peak, 2 interleaved in one warp,
9/10/11 the sweep consumer's 9x2-block MMA tile from registers with 2/4/1 warps per sub-partition,
13-15 warp-specialised DMMA + DFMA mixes. Prints TFLOP/s and ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastfp_b200 import _cabi
for k in (1, 0, 2, 9, 10, 11, 13, 14, 15):
    _cabi.fp64_peak(k, 2000)
    print(k, _cabi.fp64_peak(k, 20000))
