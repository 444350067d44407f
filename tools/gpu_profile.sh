#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, ncu launch list of the bench command and full captures of the
# dominant kernels, exported as text (the .ncu-rep files are deleted: gpurun merges at most 64 MiB back).
# Usage: bash tools/gpu_profile.sh <tag> [--with-tests]
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
if [ "${2:-}" = "--with-tests" ]; then python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -8 | tee $O/pytest_$R.log; fi
python bench.py --steps 10 --warmup 3 > $O/bench_$R.json 2> $O/bench_$R.err || { echo "bench failed"; tail -5 $O/bench_$R.err; exit 1; }
cut -c1-400 $O/bench_$R.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_$R.csv $BENCH > $O/ncu_launches_$R.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_$R.json 2> $O/bench_reference_$R.err; cut -c1-300 $O/bench_reference_$R.json
for K in k_tile_spmm k_bspmv_u k_momentum_J_cf3 k_momentum_F_thread; do
  SKIP=4; [ $K = k_tile_spmm ] && SKIP=60   # the 61st tile product of the run is a Chebyshev step (DOT = 4), the dominant instance
  ncu --set full --import-source on --clock-control none -k regex:$K -s $SKIP -c 1 -f -o /tmp/prof_$K $BENCH > $O/ncu_${K}_$R.log 2>&1
  ncu -i /tmp/prof_$K.ncu-rep --page details > $O/${R}_${K}_ncu_details.txt 2>/dev/null
  ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > $O/${R}_${K}_ncu_raw.csv 2>/dev/null
  tail -1 $O/ncu_${K}_$R.log
done
