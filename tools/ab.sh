#!/bin/bash
# A/B of step-kernel variants on the GPU box: parity subset + tools/perf_step.py for every vid2player3d_b200/lib/ab_*.so
# (built here with extra -D flags) next to the product build.  Usage: bash tools/ab.sh > gpurun_out/ab.log
cd "$(dirname "$0")/.."
D=$PWD/vid2player3d_b200/lib
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for f in $D/ab_*.so; do
  case "$f" in *ab_head.so) continue;; esac
  echo "== parity $f"; B200ENV_LIB=$f python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
done
for r in 1 2; do
  python tools/perf_step.py 8192 320
  for f in $D/ab_*.so; do B200ENV_LIB=$f python tools/perf_step.py 8192 320; done
done
