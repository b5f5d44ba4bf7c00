#!/bin/bash
# ncu --set full of the one-wave physics launch (config 2 through tools/perf_step.py, config 3 through bench.py): bash tools/ncu_tmem.sh <tag>
TAG=${1:-r2t}
cd "$(dirname "$0")/.."
export B200ENV_KERNEL=tmem
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_tmem -s 8 -c 1 -o gpurun_out/${TAG}_c2 -f \
  python tools/perf_step.py 8192 24 > gpurun_out/${TAG}_c2.log 2>&1
export B200_BENCH_PROFILE=1
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off -k regex:step_kernel_tmem -c 1 \
  -o gpurun_out/${TAG}_c3 -f python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_c3_ncu.err
ls -la gpurun_out/${TAG}_c2.ncu-rep gpurun_out/${TAG}_c3.ncu-rep
