#!/bin/bash
# launch list of the primary workload (config 3) under ncu: per-kernel GPU time of the step graph's nodes
# usage (GPU box): bash tools/federer_launches.sh <tag>   -> gpurun_out/<tag>_federer_launches.csv + a per-kernel summary on stdout
TAG=${1:-r2}
cd "$(dirname "$0")/.."
B200_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off -c 3000 --csv \
  --log-file gpurun_out/${TAG}_federer_launches.csv python bench.py --steps 6 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/${TAG}_federer_ncu_bench.json 2> gpurun_out/${TAG}_federer_ncu.err
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_federer_launches.csv")) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
names = [r[ki] for r in rows[1:]]
# one step = the launches between two consecutive physics launches of the LAST replays (steady state)
idx = [i for i, n in enumerate(names) if "step_kernel_packed" in n or "step_kernel_tmem" in n]
print("launches total", len(names), "physics launches", len(idx))
if len(idx) >= 3:
    a, b = idx[-3], idx[-2]
    seg = [r for r in rows[1:][a:b] if "FillFunctor<unsigned char>" not in r[ki]]     # the L2 flush memset between timed steps
    agg = collections.OrderedDict()
    for r in seg:
        k = r[ki][:70]
        t = float(r[vi].replace(",", ""))
        c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"one step (between two physics launches): {len(seg)} launches, {tot/1e3:.1f} us of kernel time")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {t/1e3:8.1f} us  x{c:<3d} {k}")
PY
