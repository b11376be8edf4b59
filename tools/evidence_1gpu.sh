#!/bin/bash
# round-2 evidence run on one B200 (tools/evidence_1gpu.sh): bench lines, reference arm, ncu launch lists (+ DRAM bytes), one full capture
O=gpurun_out
mkdir -p $O
echo "== default bench (C4 + secondaries + cpu baseline)"
timeout 1500 python bench.py > $O/r2_bench_c4_1gpu.json 2> $O/r2_bench_c4_1gpu.err; echo "rc=$?"; tail -c 400 $O/r2_bench_c4_1gpu.json
echo "== reference arm"
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err; echo "rc=$?"; tail -c 600 $O/r2_bench_reference_arm.json
for w in C4 C2 C3; do
  echo "== ncu launch list + DRAM bytes, $w"
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/r2_launches_$w.log 2>&1; echo "rc=$?"
done
echo "== ncu --set full, tensor sweep on C2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fp_sweep_i8_kernel -c 1 -o $O/r2_i8_c2_full -f \
  python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $O/r2_i8_c2_full.log 2>&1; echo "rc=$?"
echo "== fp64 path bench lines for comparison"
for w in C4 C2; do
  timeout 900 python bench.py --workload $w --sweep-path fp64 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/r2_bench_${w}_fp64.json 2>/dev/null; echo "rc=$?"
done
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/r2_gpu.txt
