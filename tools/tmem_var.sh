#!/bin/bash
# timing of build variants of the one-wave kernel (vid2player3d_b200/lib/ab_t_*.so), config 2 (perf_step) and config 3 (perf_federer)
cd "$(dirname "$0")/.."
D=$PWD/vid2player3d_b200/lib
export B200ENV_KERNEL=tmem
for f in $D/libb200env.so $D/ab_t_*.so; do
  B200ENV_LIB=$f timeout 200 python tools/perf_step.py 8192 96
done
for f in $D/libb200env.so $D/ab_t_*.so; do
  echo "== federer $f"; B200ENV_LIB=$f timeout 300 python tools/perf_federer.py 2>&1 | grep "physics\|step graph"
done
