#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for the bench step (1 GPU).  Run under gpurun:
#   gpurun --timeout 900 -- 'bash tools/profile_r2.sh r2e'
set -u
TAG=${1:-r2e}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c4 --no-sustained --no-sweep"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stft_kernel|mask_istft_kernel" -s 6 -c 4 \
    -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
for f in plain.log ncu_launches.log ncu_full.log; do tail -n 2 gpurun_out/${TAG}_$f; done
