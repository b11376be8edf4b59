#!/bin/bash
O=gpurun_out
mkdir -p $O
for npw in 8 16; do
  echo "== i8 bring-up + C2 timing, NPW=$npw"
  FASTFP_B200_I8_NPW=$npw timeout 600 python tools/i8_debug.py > $O/i8_debug_r2g_npw$npw.txt 2>&1; echo "rc=$?"; tail -9 $O/i8_debug_r2g_npw$npw.txt
done
echo "== ncu of the tensor sweep on C2 (NPW=8)"
FASTFP_B200_I8_NPW=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_sweep_i8_kernel -c 1 -o $O/r2g_i8_c2 -f \
  python bench.py --workload C2 --sweep-path i8 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $O/ncu_r2g.log 2>&1; echo "rc=$?"; tail -c 600 $O/ncu_r2g.log
