#!/bin/bash
# GPU run O: software-pipelined merged-operand Jacobian kernel (FB_J_KERNEL=5): parity, timing, source-level profile
set -u
O=gpurun_out
FB_J_KERNEL=5 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q --tb=short -k "(jacobian and not two_pass) or semi" 2>&1 | tail -3
for v in 2 5; do
  FB_VERBOSE=1 FB_J_KERNEL=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-variants > $O/o_bench_j$v.json 2> $O/o_bench_j$v.err
  python - <<PY
import json
d=json.load(open("$O/o_bench_j$v.json"))
print("FB_J_KERNEL=$v: step %.1f ms, phases %s, its %s, checksum %s, spmm ms %s" % (d["ms_per_step"], d["phase_ms"], d["iterations"], d["checksum"], d["roofline"]["ms_per_launch"]))
PY
done
grep flow_b200 $O/o_bench_j5.err | sort | uniq -c
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants"
FB_J_KERNEL=5 ncu --set full --import-source on --clock-control none -k regex:k_momentum_J_cf3 -s 2 -c 1 -f -o /tmp/prof_cf3 $BENCH > $O/ncu_cf3.log 2>&1
ncu -i /tmp/prof_cf3.ncu-rep --page details > $O/r2_k_momentum_J_cf3_ncu_details.txt 2>/dev/null
ncu -i /tmp/prof_cf3.ncu-rep --page raw --csv > $O/r2_k_momentum_J_cf3_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_cf3.ncu-rep --page source --csv > $O/r2_k_momentum_J_cf3_ncu_source.csv 2>/dev/null
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ" $O/r2_k_momentum_J_cf3_ncu_details.txt | head
