#!/bin/bash
# GPU run C: tile SpMM configurations (micro-benchmark) + ncu capture of selected ones.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tile.py -m gpu -q -x 2>&1 | tail -3
for c in ${CFGS:-0 1 2 3 4 5}; do
  FB_TILE_CFG=$c timeout 200 python tools/bench_spmm.py 74 50 > $O/c_spmm_cfg$c.json 2> $O/c_spmm_cfg$c.err
  python - <<PY
import json
d=json.load(open("$O/c_spmm_cfg$c.json"))
print("cfg $c", "csr %.4f ms" % d["nc3_csr"]["ms"], "tile %.4f ms" % d["nc3_tile"]["ms"], "tiles", d["nc3_tile"]["tiles"], "union/row %.2f" % d["nc3_tile"]["union_per_row"], "nc1 tile %.4f csr %.4f" % (d["nc1_tile"]["ms"], d["nc1_csr"]["ms"]), "fmt %.1fs" % d["nc3_tile"]["set_format_s"])
PY
done
for c in ${NCU_CFGS:-1}; do
  FB_TILE_CFG=$c timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_tile_spmm -s 3 -c 1 -f -o /tmp/prof_tile$c python tools/bench_spmm.py 74 5 > $O/c_ncu_tile$c.log 2>&1
  ncu -i /tmp/prof_tile$c.ncu-rep --page details > $O/c_tile_cfg${c}_ncu_details.txt 2>/dev/null
  ncu -i /tmp/prof_tile$c.ncu-rep --page raw --csv > $O/c_tile_cfg${c}_ncu_raw.csv 2>/dev/null
done
