#!/bin/bash
# GPU run B: full -m gpu suite, tile SpMM micro-benchmark, bench with the tile kernel.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -60 > $O/b_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -30 > $O/b_parity.log
timeout 300 python tools/bench_spmm.py 74 50 > $O/b_spmm.json 2> $O/b_spmm.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --variants chord > $O/b_bench.json 2> $O/b_bench.err
FB_TILE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants > $O/b_bench_notile.json 2> $O/b_bench_notile.err
tail -8 $O/b_pytest.log
cat $O/b_spmm.json
cut -c1-400 $O/b_bench.json
