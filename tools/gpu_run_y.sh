#!/bin/bash
# GPU run Y: last verification of the round -- full single-GPU test tier, smoke, default bench
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py > $O/y_bench_default.json 2> $O/y_bench_default.err; cut -c1-600 $O/y_bench_default.json
