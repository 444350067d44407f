#!/bin/bash
# GPU run P: cf3 with the lane = (n, k) phase A as the 3D default, tile-ordered r / z / dinv in the Chebyshev epilogue,
# Stokes timing (FB_VERBOSE)
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_tile.py -m gpu -q --tb=short 2>&1 | tail -3
FB_J_KERNEL=5 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q --tb=short -k "(jacobian and not two_pass) or semi" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -E "fixture|passed|failed" | cut -c1-420
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants > $O/p_bench.json 2> $O/p_bench.err
python - <<PY
import json
d=json.load(open("$O/p_bench.json"))
print("step %.1f ms, phases %s, its %s, checksum %s, spmm ms %s" % (d["ms_per_step"], d["phase_ms"], d["iterations"], d["checksum"], d["roofline"]["ms_per_launch"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tile_spmm|k_momentum_J|k_momentum_F" --csv --log-file $O/p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants > /dev/null 2>&1
python tools/summarise_profiles.py $O/p_launches.csv | head -12
FB_VERBOSE=1 timeout 300 python tools/run_configs.py karman --steps 50 2>&1 | grep -E "stokes|config" | cut -c1-600
