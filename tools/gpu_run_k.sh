#!/bin/bash
# GPU run K: register-bound variants of the Jacobian element kernel; AMG on S (variants test, config 2).
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q --tb=short 2>&1 | tail -6
timeout 600 python bench.py --config 2 > $O/k_config2.json 2> $O/k_config2.err
python - <<PY
import json
try:
    d=json.load(open("$O/k_config2.json")); x=d["details"]
    print("config 2: %.2f steps/s e2e" % d["value"], {k:x[k] for k in x if k not in ("config",)})
except Exception as e:
    print("config 2 failed", e); print(open("$O/k_config2.err").read()[-800:])
PY
for m in 3 4 5 6; do
  FB_J_MINB=$m timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --no-variants > $O/k_bench_minb$m.json 2> $O/k_bench_minb$m.err
  python - <<PY
import json
d=json.load(open("$O/k_bench_minb$m.json"))
print("MINB $m: step %.1f ms, assembly_J %.2f ms per step (%.1f assemblies)" % (d["ms_per_step"], d["phase_ms"]["assembly_J"], d["iterations"]["jacobian_assemblies"]))
PY
done
