#!/bin/bash
# GPU run W (8 GPUs): final n = 74 line at N = 8 with the graph-replayed V-cycle
set -u
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-variants > $O/w_bench_n74_N8.json 2> $O/w_bench_n74_N8.err
python - <<PY
import json
d=json.loads([l for l in open("$O/w_bench_n74_N8.json") if l.startswith("{")][0])
print("N=8", "steps/s %.3f ms %.1f" % (d["value"], d["ms_per_step"]), d["iterations"], d["phase_ms"], d["checksum"], "e2e", d["e2e"]["value"], d.get("comm"))
PY
