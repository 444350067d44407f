#!/bin/bash
# GPU run M: merged-operand Jacobian kernel (FB_J_KERNEL=5) parity + timing against the default; Chebyshev epilogue with
# prefetched row operands; host-side profile of configs 3 and 4.
set -u
O=gpurun_out
mkdir -p $O
FB_J_KERNEL=5 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q --tb=short -k "jacobian or semi" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_heat_stokes.py -m gpu -q --tb=short -k "stokes" 2>&1 | tail -4
for v in 2 5; do
  FB_J_KERNEL=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-variants > $O/m_bench_j$v.json 2> $O/m_bench_j$v.err
  python - <<PY
import json
d=json.load(open("$O/m_bench_j$v.json"))
print("FB_J_KERNEL=$v: step %.1f ms, phases %s, its %s, checksum %s, spmm ms %s" % (d["ms_per_step"], d["phase_ms"], d["iterations"]["momentum_krylov"], d["checksum"], d["roofline"]["ms_per_launch"]))
PY
done
timeout 300 python -m cProfile -s cumtime tools/run_configs.py boussinesq --steps 3 2>&1 | head -60 > $O/m_prof_boussinesq.txt
timeout 300 python -m cProfile -s cumtime tools/run_configs.py karman --steps 20 2>&1 | head -50 > $O/m_prof_karman.txt
timeout 300 python tools/sweep_cheb_degree.py 4 8 12 16 24 2>&1 | tail -6 | tee $O/m_cheb_degree_config2.jsonl | cut -c1-400
