#!/bin/bash
# timing of build variants of the one-wave kernel (vid2player3d_b200/lib/ab_t_*.so): config 2 (perf_step), config 3 (perf_federer, bench value)
cd "$(dirname "$0")/.."
D=$PWD/vid2player3d_b200/lib
for f in $D/libb200env.so $D/ab_t_*.so; do
  B200ENV_LIB=$f timeout 200 python tools/perf_step.py 8192 96
done
for f in $D/libb200env.so $D/ab_t_*.so; do
  echo "== federer $f"; B200ENV_LIB=$f timeout 300 python tools/perf_federer.py 2>&1 | grep "physics\|step graph\|step + reset"
  B200ENV_LIB=$f timeout 300 python bench.py --steps 20 --warmup 5 --legs none --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench', {k: d[k] for k in ('value', 'ms_per_step', 'value_hot_l2_back_to_back')}, d['roofline']['dominant_kernel']['ms'])"
done
