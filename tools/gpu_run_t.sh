#!/bin/bash
# GPU run T (4 GPUs): final strong-scaling line of round 2 at N = 4
set -u
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-variants > $O/t_bench_n74_N4.json 2> $O/t_bench_n74_N4.err
python - <<PY
import json
d=json.loads([l for l in open("$O/t_bench_n74_N4.json") if l.startswith("{")][0])
print("N=4", "steps/s %.3f ms %.1f" % (d["value"], d["ms_per_step"]), d["iterations"], d["phase_ms"], d["checksum"], "e2e", d["e2e"]["value"], d.get("comm"))
PY
