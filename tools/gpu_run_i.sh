#!/bin/bash
# GPU run I: full gpu tests with complete failure output, bench without the CPU leg.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_multi.py > $O/i_pytest.log 2>&1
tail -4 $O/i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/i_bench.json 2> $O/i_bench.err
python - <<PY
import json
d=json.load(open("$O/i_bench.json"))
print(d["value"], d["ms_per_step"], d["iterations"], d["phase_ms"], d["e2e"])
print({k:(v.get("ms_per_step"), v.get("error")) for k,v in d["variants"].items()})
r=d["roofline"]; print(r["kernel"][:40], r["ms_per_launch"], r["launches_per_step"], r["share_of_step"])
PY
