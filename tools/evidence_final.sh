#!/bin/bash
# last check of a round on one B200: GPU parity suite, compute-sanitizer memcheck of a small run, default bench line
O=gpurun_out
mkdir -p $O
echo "== GPU parity suite"
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2_pytest_gpu.log
echo "== compute-sanitizer memcheck (tensor sweep + fp64 sweep + nmfp + get_xCy, small)"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r2_sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -6 $O/r2_sanitizer_memcheck.log
echo "== default bench"
timeout 1500 python bench.py > $O/r2_bench_c4_1gpu.json 2> $O/r2_bench_c4_1gpu.err; echo "rc=$?"; tail -c 300 $O/r2_bench_c4_1gpu.json
