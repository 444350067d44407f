#!/usr/bin/env python
"""Turn the ncu launch list of a bench.py run (gpu__time_duration.sum per launch, --csv) into the per-kernel share
table kept under profiles/.  Usage: python tools/summarise_profiles.py gpurun_out/launches_bench_r1.csv > profiles/r1_launch_shares.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        k = r[ki].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        k = k.split("(")[0] if not k.startswith("(") else k
        k = k.replace("(int)", "")
        agg[k][0] += 1
        agg[k][1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.1f ms of kernel time (ncu per-launch times are cold-cache and serialised:" % (path, sum(v[0] for v in agg.values()), tot))
    print("# compare SHARES with the CUDA-event phase times of bench.py, not absolutes)")
    print("%-64s %7s %11s %7s %10s" % ("kernel", "calls", "total ms", "share", "avg ms"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 5e-4:
            continue
        print("%-64s %7d %11.3f %6.1f%% %10.4f" % (k[:64], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1])
