#!/bin/bash
# ncu --set full captures of the primary workload's dominant kernels (GPU box): bash tools/ncu_capture.sh <tag>
#   gpurun_out/<tag>_step.ncu-rep   step_kernel_packed<split> of config 3 (12 substeps + ball), first launch of the timed loop
#   gpurun_out/<tag>_gemm.ncu-rep   linear_kernel launches of the same step (policy layers + mixture-of-experts layers)
TAG=${1:-r2}
cd "$(dirname "$0")/.."
export B200_BENCH_PROFILE=1
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off -k "regex:step_kernel_(packed|tmem)" -c 1 \
  -o gpurun_out/${TAG}_step -f python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_step_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off -k regex:linear_kernel -c 9 \
  -o gpurun_out/${TAG}_gemm -f python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_gemm_ncu.err
ls -la gpurun_out/${TAG}_step.ncu-rep gpurun_out/${TAG}_gemm.ncu-rep
