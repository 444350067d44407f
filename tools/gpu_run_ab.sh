#!/bin/bash
# GPU run AB: warm start of the first update from the extrapolated rate of the last two steps (FB_WARM_DELTA=2)
set -u
O=gpurun_out
for g in 2; do
  FB_WARM_DELTA=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-variants > $O/ab_bench_w$g.json 2> $O/ab_bench_w$g.err
  python - <<PY
import json
d=json.load(open("$O/ab_bench_w$g.json"))
print("FB_WARM_DELTA=$g: step %.1f ms, momentum_solve %.2f ms, its %s, checksum %s, newton residuals %s" % (d["ms_per_step"], d["phase_ms"]["momentum_solve"], d["iterations"], d["checksum"], d["newton_residuals_last_step"]))
PY
done
FB_WARM_DELTA=2 timeout 900 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -E "fixture|passed|failed" | cut -c1-330
