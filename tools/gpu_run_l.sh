#!/bin/bash
# GPU run L: column-per-lane Jacobian kernel (FB_J_KERNEL=3): parity + timing against the default.
set -u
O=gpurun_out
mkdir -p $O
FB_J_KERNEL=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "jacobian" 2>&1 | tail -4
for v in 2 3; do
  FB_J_KERNEL=$v timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --no-variants > $O/l_bench_j$v.json 2> $O/l_bench_j$v.err
  python - <<PY
import json
d=json.load(open("$O/l_bench_j$v.json"))
print("FB_J_KERNEL=$v: step %.1f ms, assembly_J %.2f ms per step, its %s, checksum %s" % (d["ms_per_step"], d["phase_ms"]["assembly_J"], d["iterations"]["momentum_krylov"], d["checksum"]))
PY
done
