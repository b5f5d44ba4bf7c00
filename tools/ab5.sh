cd /root/repo
D=$PWD/vid2player3d_b200/lib
timeout 600 python -m pytest tests/test_gpu_v2p.py tests/test_gpu_tmem.py -m gpu -q 2>&1 | tail -4
for r in 1 2; do
  timeout 200 python tools/perf_step.py 8192 96
  B200ENV_LIB=$D/ab_noffma2.so timeout 200 python tools/perf_step.py 8192 96
done
for f in $D/libb200env.so $D/ab_noffma2.so; do
  echo "== $f"; B200ENV_LIB=$f timeout 300 python tools/perf_federer.py 2>&1 | grep "physics\|step graph\|step + reset"
  B200ENV_LIB=$f timeout 300 python bench.py --steps 20 --warmup 5 --legs none --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench', {k: d[k] for k in ('value', 'ms_per_step', 'value_hot_l2_back_to_back', 'gpu_launches')}, d['e2e']['value'], d['roofline']['dominant_kernel']['ms'])"
done
