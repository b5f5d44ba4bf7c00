#!/bin/bash
# the round's final evidence on a GPU box: bash tools/final_captures.sh <tag>
TAG=${1:-r2ac}
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
bash tools/ncu_tmem.sh ${TAG} > gpurun_out/${TAG}_ncu.log 2>&1
bash tools/federer_launches.sh ${TAG} > gpurun_out/${TAG}_federer_launches.txt 2>&1
timeout 200 python tools/step_timeline.py > gpurun_out/${TAG}_timeline.log 2>&1
timeout 300 python tools/perf_federer.py > gpurun_out/${TAG}_federer_components.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/${TAG}_sanitizer.log 2>&1
echo "=== racecheck" >> gpurun_out/${TAG}_sanitizer.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py >> gpurun_out/${TAG}_sanitizer.log 2>&1
tail -c 300 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_sanitizer.log; tail -3 gpurun_out/${TAG}_timeline.log; cat gpurun_out/${TAG}_federer_components.log | tail -12
