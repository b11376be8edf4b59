#!/bin/bash
# Round-2 GPU session B: the re-written producers (lockstep sincos, 8 vs 16 producer warps), parity + timing + ncu
O=gpurun_out
mkdir -p $O
echo "== A-from-TMEM probe"; (cd tools/probes && timeout 60 ./umma_i8_ts_probe 2000) > $O/ts_probe_r2.txt 2>&1; cat $O/ts_probe_r2.txt
for npw in 8 16; do
  echo "== i8 bring-up, $npw producer warps"
  FASTFP_B200_I8_NPW=$npw timeout 300 python tools/i8_debug.py > $O/i8_debug_npw$npw.txt 2>&1; echo "rc=$?"; tail -9 $O/i8_debug_npw$npw.txt
done
echo "== i8 bring-up, A planes in tensor memory (TS form), 8 producer warps"
FASTFP_B200_I8_TS=1 timeout 300 python tools/i8_debug.py > $O/i8_debug_ts.txt 2>&1; echo "rc=$?"; tail -9 $O/i8_debug_ts.txt
if ! grep -q "ALL OK" $O/i8_debug_npw8.txt; then echo "i8 bring-up failed: stopping"; exit 0; fi
echo "== GPU parity suite on the tensor path"
FASTFP_B200_PATH=prefer-i8 timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_r2_i8.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_r2_i8.log
echo "== ncu full capture of the tensor sweep kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fp_sweep_i8_kernel -s 2 -c 1 -o $O/prof_r2_i8b env FASTFP_B200_PATH=prefer-i8 python tools/prof_sweep.py C2 2048 3 > $O/prof_r2_i8b.log 2>&1; echo "rc=$?"; tail -3 $O/prof_r2_i8b.log
