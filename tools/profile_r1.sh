#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for the bench step (1 GPU).  Run under gpurun.
set -u
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft -s 6 -c 2 \
    -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
for f in plain.log ncu_launches.log ncu_full.log; do tail -n 3 gpurun_out/$f; done
