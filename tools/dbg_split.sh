#!/bin/bash
# step times of the C2 sweep with parts of the kernel switched off (FASTFP_DBG bits: 1 = producers skip the
# sincos math, 2 = consumers skip the MMAs, 4 = no level-2 flush): 3 = the pipeline skeleton alone
for d in ${@:-0 1 2 3}; do
  FASTFP_DBG=$d timeout 100 python bench.py --no-cpu-baseline --steps 5 2>/dev/null > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('dbg', $d, round(d['ms_per_step'],3))"
done
