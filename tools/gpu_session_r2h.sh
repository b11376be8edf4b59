#!/bin/bash
O=gpurun_out
mkdir -p $O
for npw in 16 8; do
  for w in C4 C3; do
    echo "== bench $w NPW=$npw"
    FASTFP_B200_I8_NPW=$npw timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/bench_r2h_${w}_npw$npw.json 2> $O/bench_r2h_${w}_npw$npw.err; echo "rc=$?"
    python - <<PY
import json
d=json.loads(open("$O/bench_r2h_${w}_npw$npw.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"].get("sweep_kernel"), d.get("checks_failed"))
PY
  done
done
echo "== GPU parity suite (auto = tensor kernel)"
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_r2h.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_r2h.log
