#!/bin/bash
# GPU run Z: residual kernel with the test-node-independent part hoisted: parity tests, timing
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_reference_tests.py -m gpu -q --tb=short 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -E "fixture|passed|failed" | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants > $O/z_bench.json 2> $O/z_bench.err
python - <<PY
import json
d=json.load(open("$O/z_bench.json"))
print("step %.1f ms (%.3f steps/s), e2e %s, phases %s, its %s, checksum %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["phase_ms"], d["iterations"], d["checksum"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_momentum_F" --csv --log-file $O/z_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants > /dev/null 2>&1
python tools/summarise_profiles.py $O/z_launches.csv | tail -3
