#!/bin/bash
# step times of the C2 sweep with parts of the kernel switched off (FASTFP_DBG bits: 1 = producers skip the
# sincos math, 2 = consumers skip the MMAs, 4 = no level-2 flush): 3 = the pipeline skeleton alone.
# The switches exist only in a developer build: FASTFP_B200_NVCC_FLAGS=-DFFP_DEBUG_SWITCHES python -m
# fastfp_b200.build --force (the shipped library has no such run-time switch; bench.py refuses FASTFP_DBG).
for d in ${@:-0 1 2 3}; do
  FASTFP_DBG=$d FASTFP_BENCH_ALLOW_DBG=1 timeout 100 python bench.py --workload C2 --no-cpu-baseline --steps 5 2>/dev/null > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('dbg', $d, round(d['ms_per_step'],3))"
done
