#!/bin/bash
# one B200: full GPU suite on the final library, then the sweep time against the basis width on both kernels
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2_pytest_gpu.log
timeout 900 python tools/time_widths.py > $O/r2_time_widths.txt 2>&1; echo "rc=$?"; cat $O/r2_time_widths.txt
