#!/bin/bash
# GPU run E (2 GPUs): partitioned-run tests and a 2-GPU bench line.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > $O/e_pytest_multi.log
tail -5 $O/e_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-variants > $O/e_bench_n2.json 2> $O/e_bench_n2.err
tail -3 $O/e_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("$O/e_bench_n2.json") if l.startswith("{")][0])
print(d["value"], d["ms_per_step"], d["iterations"], d["phase_ms"], d["checksum"], d["e2e"])
PY
