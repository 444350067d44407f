#!/bin/bash
# GPU run AC: last line of the round (warm-started first update and correction solve): parity at defaults, variants, bench
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -E "fixture|passed|failed" | cut -c1-330 | tee $O/parity_r2.log
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_reference_tests.py tests/test_gpu_drivers.py -m gpu -q --tb=short 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > $O/bench_r2.json 2> $O/bench_r2.err; cut -c1-200 $O/bench_r2.json
python - <<PY
import json
d=json.load(open("$O/bench_r2.json"))
print("step %.1f ms (%.3f steps/s), e2e %s, phases %s, its %s, checksum %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["phase_ms"], d["iterations"], d["checksum"]))
PY
