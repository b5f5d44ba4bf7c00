#!/bin/bash
# the last measurements of the round on a GPU box: bash tools/final_bench.sh <tag>
TAG=${1:-r2ae}
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 200 gpurun_out/${TAG}_bench.json
timeout 300 python tools/perf_federer.py > gpurun_out/${TAG}_federer_components.log 2>&1; tail -12 gpurun_out/${TAG}_federer_components.log
timeout 200 python tools/perf_step.py 8192 96
bash tools/federer_launches.sh ${TAG} > gpurun_out/${TAG}_federer_launches.txt 2>&1; head -12 gpurun_out/${TAG}_federer_launches.txt
[ -e vid2player3d_b200/lib/ab_prof.so ] && B200ENV_LIB=$PWD/vid2player3d_b200/lib/ab_prof.so timeout 200 python tools/pt_prof.py amass 5 30 > gpurun_out/${TAG}_pt_prof_amass.log 2>&1 && cat gpurun_out/${TAG}_pt_prof_amass.log
