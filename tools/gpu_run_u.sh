#!/bin/bash
# GPU run U: CUDA-graph replay of the AMG V-cycle: tests that use the hierarchy, bench with and without, config 4, smoke
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_heat_stokes.py tests/test_gpu_reference_tests.py tests/test_gpu_drivers.py -m gpu -q --tb=short 2>&1 | tail -3
for g in 0 1; do
  FB_AMG_GRAPH=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants > $O/u_bench_g$g.json 2> $O/u_bench_g$g.err
  python - <<PY
import json
d=json.load(open("$O/u_bench_g$g.json"))
print("FB_AMG_GRAPH=$g: step %.1f ms, phases %s, its %s, checksum %s, launches %s" % (d["ms_per_step"], d["phase_ms"], d["iterations"], d["checksum"], d["gpu_launches"]))
PY
done
timeout 300 python tools/run_configs.py boussinesq karman --steps 20 2>&1 | grep config | cut -c1-900
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
