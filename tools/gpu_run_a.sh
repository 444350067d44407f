#!/bin/bash
# GPU run A (round 2): full -m gpu suite, default bench (reference-exact Newton) with the chord variant beside it,
# and the experimental coordinate node numbering timed for the first time.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $O/a_gpu.txt; nproc >> $O/a_gpu.txt; free -g | head -2 >> $O/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -40 > $O/a_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity_default.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 > $O/a_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --variants chord > $O/a_bench.json 2> $O/a_bench.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants --node-order lexicographic > $O/a_bench_lex.json 2> $O/a_bench_lex.err
tail -5 $O/a_pytest.log
cut -c1-600 $O/a_bench.json
