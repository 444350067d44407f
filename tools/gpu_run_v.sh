#!/bin/bash
# GPU run V: lanes per row of the AMG kernels (FB_AMG_LANES=0: first thresholds), Stokes timing
set -u
O=gpurun_out
for g in 0 1; do
  FB_AMG_LANES=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants > $O/v_bench_l$g.json 2> $O/v_bench_l$g.err
  python - <<PY
import json
d=json.load(open("$O/v_bench_l$g.json"))
print("FB_AMG_LANES=$g: step %.1f ms, pressure %.3f ms, its %s, checksum %s" % (d["ms_per_step"], d["phase_ms"]["pressure"], d["iterations"]["pressure_cg"], d["checksum"]))
PY
done
FB_VERBOSE=1 timeout 300 python tools/run_configs.py karman --steps 20 2>&1 | grep -E "stokes:|config" | cut -c1-700
FB_VERBOSE=1 timeout 300 python tools/run_configs.py boussinesq karman --steps 5 2>&1 | grep -E "stokes:|config" | cut -c1-700
