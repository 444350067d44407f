#!/bin/bash
# GPU run R (2 GPUs): partitioned-run tests and the 2-GPU bench line after the kernel changes of the last session
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-variants > $O/r_bench_n2.json 2> $O/r_bench_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r_bench_n2.json") if l.startswith("{")][0])
    print("N=2", d["value"], d["ms_per_step"], d["iterations"], d["phase_ms"], d["checksum"], d.get("comm"), d.get("e2e"))
except Exception as e:
    print("failed", e); print(open("$O/r_bench_n2.err").read()[-1500:])
PY
