#!/bin/bash
# ncu --set full of the one-wave physics launch: bash tools/ncu_tmem.sh <tag>
#   <tag>_c2   config 2 through tools/perf_step.py, launch 50 = 30 steps after a reset (most humanoids on the ground: contact-heavy)
#   <tag>_c3   config 3 through bench.py, first launch of the timed loop (5 steps after the reset: standing / falling humanoids)
TAG=${1:-r2t}
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel_tmem -s 50 -c 1 -o gpurun_out/${TAG}_c2 -f \
  python tools/perf_step.py 8192 40 > gpurun_out/${TAG}_c2.log 2>&1
export B200_BENCH_PROFILE=1
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off -k regex:step_kernel_tmem -c 1 \
  -o gpurun_out/${TAG}_c3 -f python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_c3_ncu.err
ls -la gpurun_out/${TAG}_c2.ncu-rep gpurun_out/${TAG}_c3.ncu-rep
