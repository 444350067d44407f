#!/bin/bash
# GPU run Q: block cache behind DBuf, Stokes with separate work vectors: full single-GPU test tier, configs 3 and 4, bench
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -5
FB_VERBOSE=1 timeout 300 python tools/run_configs.py karman boussinesq cavity2d --steps 20 2>&1 | grep -E "stokes:|config" | cut -c1-900 | tee $O/q_configs.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants > $O/q_bench.json 2> $O/q_bench.err
python - <<PY
import json
d=json.load(open("$O/q_bench.json"))
print("step %.1f ms, e2e %s, phases %s, its %s, checksum %s, spmm ms %s" % (d["ms_per_step"], d["e2e"], d["phase_ms"], d["iterations"], d["checksum"], d["roofline"]["ms_per_launch"]))
PY
