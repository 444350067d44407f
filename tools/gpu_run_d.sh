#!/bin/bash
# GPU run D: tile SpMM (TMA configs) after the half-warp scheduling fix, full gpu tests, bench with Chebyshev vs CG inner.
set -u
O=gpurun_out
mkdir -p $O
for c in 0 5; do
  FB_TILE_CFG=$c timeout 200 python tools/bench_spmm.py 74 50 > $O/d_spmm_cfg$c.json 2> $O/d_spmm_cfg$c.err
  python - <<PY
import json
d=json.load(open("$O/d_spmm_cfg$c.json"))
print("cfg $c", "csr %.4f ms" % d["nc3_csr"]["ms"], "tile %.4f ms" % d["nc3_tile"]["ms"], "tiles", d["nc3_tile"]["tiles"], "union/row %.2f" % d["nc3_tile"]["union_per_row"], "nc1 tile %.4f csr %.4f" % (d["nc1_tile"]["ms"], d["nc1_csr"]["ms"]))
PY
done
FB_TILE_CFG=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_tile_spmm -s 3 -c 1 -f -o /tmp/prof_tile0 python tools/bench_spmm.py 74 5 > $O/d_ncu_tile0.log 2>&1
ncu -i /tmp/prof_tile0.ncu-rep --page details > $O/d_tile_cfg0_ncu_details.txt 2>/dev/null
ncu -i /tmp/prof_tile0.ncu-rep --page raw --csv > $O/d_tile_cfg0_ncu_raw.csv 2>/dev/null
export FB_TILE_CFG=0
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -30 > $O/d_pytest.log
tail -5 $O/d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --variants "" > $O/d_bench.json 2> $O/d_bench.err
cut -c1-300 $O/d_bench.json
python - <<PY
import json
d=json.load(open("$O/d_bench.json"))
print(d["ms_per_step"], d["iterations"], d["phase_ms"], d["newton_residuals_last_step"], d["checksum"])
PY
