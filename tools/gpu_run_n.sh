#!/bin/bash
# GPU run N: ncu --set full of the merged-operand Jacobian kernel (FB_J_KERNEL=5)
set -u
O=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants"
FB_J_KERNEL=5 ncu --set full --import-source on --clock-control none -k regex:k_momentum_J_cf3 -s 2 -c 1 -f -o /tmp/prof_cf3 $BENCH > $O/ncu_cf3.log 2>&1
ncu -i /tmp/prof_cf3.ncu-rep --page details > $O/r2_k_momentum_J_cf3_ncu_details.txt 2>/dev/null
ncu -i /tmp/prof_cf3.ncu-rep --page raw --csv > $O/r2_k_momentum_J_cf3_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_cf3.ncu-rep --page source --csv > $O/r2_k_momentum_J_cf3_ncu_source.csv 2>/dev/null
tail -2 $O/ncu_cf3.log; ls -la $O/r2_k_momentum_J_cf3*
timeout 600 python -m pytest tests/test_gpu_heat_stokes.py -m gpu -q --tb=short -k "stokes" 2>&1 | tail -4
timeout 300 python tools/run_configs.py karman --steps 100 2>&1 | tail -1 | cut -c1-700 | tee $O/n_karman.json
FB_VERBOSE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-variants 2>&1 >$O/n_bench.json | grep flow_b200 | sort | uniq -c
python -c "
import json;d=json.load(open('$O/n_bench.json'));print(d['ms_per_step'],d['phase_ms'],d['iterations'])"
FB_VERBOSE=1 timeout 300 python tools/sweep_cheb_degree.py 0 2>&1 | grep -E "flow_b200|cheb" | sort | uniq -c | cut -c1-400
