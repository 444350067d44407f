#!/bin/bash
# GPU run S (8 GPUs): final strong-scaling lines of round 2 -- n = 74 at N = 8, the 80 M-dof cavity (n = 147) at N = 8
set -u
O=gpurun_out
mkdir -p $O
run() { # N n tag extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 --cube-n $2 --steps ${STEPS:-5} --warmup 3 --no-cpu --no-variants $4 > $O/s_bench_$3.json 2> $O/s_bench_$3.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/s_bench_$3.json") if l.startswith("{")][0])
    print("$3", "steps/s %.3f ms %.1f" % (d["value"], d["ms_per_step"]), d["iterations"], d["phase_ms"], d["checksum"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "setup %.0fs" % d["setup_s"], d.get("comm"))
except Exception as e:
    print("$3 failed", e); print(open("$O/s_bench_$3.err").read()[-1500:])
PY
}
run 8 74 n74_N8 ""
run 8 147 n147_N8 "--no-e2e"
