#!/bin/bash
# Round-2 GPU session: bring-up of the tensor sweep, then (only if it is correct) the parity suite on it, the bench
# line and the ncu captures. Every step has its own timeout; logs land in gpurun_out/.
O=gpurun_out
mkdir -p $O
echo "== i8 bring-up"; timeout 300 python tools/i8_debug.py > $O/i8_debug.txt 2>&1; echo "rc=$?"; tail -12 $O/i8_debug.txt
if ! grep -q "ALL OK" $O/i8_debug.txt; then echo "i8 bring-up failed: stopping"; exit 0; fi
echo "== GPU parity suite on the tensor path"
FASTFP_B200_PATH=prefer-i8 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_r2_i8.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_r2_i8.log
echo "== bench C4 on the tensor path"
timeout 600 python bench.py --steps 5 --warmup 3 --sweep-path prefer-i8 > $O/bench_r2_c4_1gpu_i8.json 2> $O/bench_r2_c4_1gpu_i8.err; echo "rc=$?"; head -c 600 $O/bench_r2_c4_1gpu_i8.json; tail -2 $O/bench_r2_c4_1gpu_i8.err
echo "== ncu launch list (C2, tensor path)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_r2_c2_i8.csv python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline --sweep-path prefer-i8 > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full capture of the tensor sweep kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fp_sweep_i8_kernel -s 2 -c 1 -o $O/prof_r2_i8 env FASTFP_B200_PATH=prefer-i8 python tools/prof_sweep.py C2 2048 3 > $O/prof_r2_i8.log 2>&1; echo "rc=$?"; tail -3 $O/prof_r2_i8.log
