#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, ncu launch list of the bench command and full captures of the
# dominant kernels.  Everything lands in gpurun_out/; summaries are copied to profiles/ by tools/summarise_profiles.py.
set -u
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > $O/bench_$R.json 2> $O/bench_$R.err || { echo "bench failed"; tail -5 $O/bench_$R.err; exit 1; }
cut -c1-600 $O/bench_$R.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_$R.csv $BENCH > $O/ncu_launches_$R.log 2>&1
for K in k_bspmv_u k_spmm_u k_momentum_J k_momentum_F_thread; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o $O/prof_${K}_$R $BENCH > $O/ncu_${K}_$R.log 2>&1
  tail -1 $O/ncu_${K}_$R.log
done
