#!/bin/bash
# GPU run E (2 GPUs): partitioned-run tests and 2-GPU bench lines (rank-local vs global Chebyshev preconditioner).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -8 > $O/e_pytest_multi.log
tail -4 $O/e_pytest_multi.log
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-variants --no-e2e $2 > $O/e_bench_n2_$1.json 2> $O/e_bench_n2_$1.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/e_bench_n2_$1.json") if l.startswith("{")][0])
    print("$1", d["value"], d["ms_per_step"], d["iterations"], d["phase_ms"], d["checksum"])
except Exception as e:
    print("$1 failed", e); print(open("$O/e_bench_n2_$1.err").read()[-1200:])
PY
}
run local ""
FB_INNER_LOCAL=0 run global ""
