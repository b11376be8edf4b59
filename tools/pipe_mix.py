import sys; sys.path.insert(0,'/root/repo')
from fastfp_b200 import _cabi
for k in (1,0,2,13,14,15):
    _cabi.fp64_peak(k, 2000)
    print(k, _cabi.fp64_peak(k, 20000))
