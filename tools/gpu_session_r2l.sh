#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python tools/i8_debug.py | tail -4
echo "== GPU parity suite"
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_r2l.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_r2l.log
for w in C2 C3; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > $O/bench_r2l_$w.json 2>/dev/null; python -c "
import json;d=json.loads(open('$O/bench_r2l_$w.json').read().strip().splitlines()[-1]);print('$w',d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['run'].get('content_hash_ms_per_call'))"; done
