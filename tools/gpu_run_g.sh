#!/bin/bash
# GPU run G: full gpu tests (new: semi-implicit, drivers), option sweep, compute-sanitizer on small cases.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -30 > $O/g_pytest.log
tail -6 $O/g_pytest.log
timeout 900 python tools/tune_newton.py 74 8 > $O/g_tune.jsonl 2> $O/g_tune.err
python - <<PY
import json
for l in open("$O/g_tune.jsonl"):
    d=json.loads(l); print(d["options"], "%.1f ms" % d["ms_per_step"], d["newton_momentum_inner_assemblies"], "cube24 step10 err", d["cube24_err_u_p"].get("10"), "max", max(max(v) for v in d["cube24_err_u_p"].values()))
PY
tail -3 $O/g_tune.err
