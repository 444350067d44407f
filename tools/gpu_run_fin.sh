#!/bin/bash
# Last evidence refresh of round 2 (after the residual-kernel change): bench line, launch list, full capture of the
# residual kernel; the captures of the other three kernels (tools/gpu_profile.sh, same session) are unchanged code.
set -u
O=gpurun_out
R=r2
python bench.py --steps 10 --warmup 3 > $O/bench_$R.json 2> $O/bench_$R.err || { echo "bench failed"; tail -5 $O/bench_$R.err; exit 1; }
cut -c1-300 $O/bench_$R.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_$R.csv $BENCH > $O/ncu_launches_$R.log 2>&1
K=k_momentum_F_thread
ncu --set full --import-source on --clock-control none -k regex:$K -s 4 -c 1 -f -o /tmp/prof_$K $BENCH > $O/ncu_${K}_$R.log 2>&1
ncu -i /tmp/prof_$K.ncu-rep --page details > $O/${R}_${K}_ncu_details.txt 2>/dev/null
ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > $O/${R}_${K}_ncu_raw.csv 2>/dev/null
tail -1 $O/ncu_${K}_$R.log
