"""fp64-pipe probes (fastfp_fp64_peak kinds): 1 DMMA peak, 0 DFMA peak, 2 interleaved in one warp,
9/10/11 the sweep consumer's 9x2-block MMA tile from registers with 2/4/1 warps per sub-partition,
13-15 warp-specialised DMMA + DFMA mixes, 16 legacy INT8 mma.sync (T(FL)OP/s = 2 x MAC/s). Prints TFLOP/s and ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastfp_b200 import _cabi
for k in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (1, 0, 2, 9, 10, 11, 13, 14, 15, 16):
    _cabi.fp64_peak(k, 2000)
    print(k, _cabi.fp64_peak(k, 20000))
