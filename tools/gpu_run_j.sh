#!/bin/bash
# GPU run J: compute-sanitizer (memcheck, racecheck, synccheck) on the small workload; configs 2-4; gpu tests.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "semi_implicit or two_pass" 2>&1 | tail -5
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > $O/j_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case: done" $O/j_sanitizer_$tool.log | tail -3
done
for c in 2 3 4; do
  timeout 900 python bench.py --config $c > $O/j_config$c.json 2> $O/j_config$c.err
  python - <<PY
import json
try:
    d=json.load(open("$O/j_config$c.json")); x=d["details"]
    print("config $c: %.2f steps/s e2e" % d["value"], {k:x[k] for k in x if k not in ("config",)})
except Exception as e:
    print("config $c failed", e); print(open("$O/j_config$c.err").read()[-800:])
PY
done
