#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== overlap probe (fp64 / integer / fp32 / conversions next to tcgen05 MMAs)"
(cd tools/probes && timeout 120 ./umma_fp64_overlap_probe 3000) > $O/overlap_probe_r2.txt 2>&1; cat $O/overlap_probe_r2.txt
echo "== GPU parity suite, tensor kernel wherever the pack can take it"
FASTFP_B200_TEST_PREFER_I8=1 FASTFP_B200_PATH=prefer-i8 timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_r2_i8.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_r2_i8.log
grep -n "worst |got" $O/pytest_r2_i8.log | head -5
