#!/bin/bash
# A/B of the one-wave physics kernel (step_kernel_tmem, csrc/packed_t.cuh; the default) against step_kernel_packed on a GPU box:
# equality tests first (exact in a -fmad=false build if vid2player3d_b200/lib/ab_nofma.so exists, rounding-level in the product build),
# then tools/perf_step.py (config 2), tools/perf_federer.py (config 3) and the bench line of the primary workload for both.
cd "$(dirname "$0")/.."
D=$PWD/vid2player3d_b200/lib
[ -e $D/ab_nofma.so ] && { echo "== exact equality, -fmad=false build"; B200ENV_LIB=$D/ab_nofma.so B200_EXPECT_EXACT=1 timeout 600 python -m pytest tests/test_gpu_tmem.py -q 2>&1 | tail -4; }
echo "== product build"; timeout 600 python -m pytest tests/test_gpu_tmem.py -q 2>&1 | tail -4
for r in 1 2; do
  B200ENV_KERNEL=packed timeout 200 python tools/perf_step.py 8192 96
  timeout 200 python tools/perf_step.py 8192 96
done
for k in packed tmem; do
  echo "== federer B200ENV_KERNEL=$k"; B200ENV_KERNEL=$k timeout 300 python tools/perf_federer.py 2>&1 | grep "step\|physics"
done
for k in packed tmem; do
  echo "== bench B200ENV_KERNEL=$k"; B200ENV_KERNEL=$k timeout 300 python bench.py --steps 20 --warmup 5 --legs none --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'value_hot_l2_back_to_back')}, d['e2e']['value'], d['roofline']['dominant_kernel']['ms'])"
done
