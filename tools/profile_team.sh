set -u
CMD="python tools/kbench.py 1024 256 1024 48000"
$CMD > gpurun_out/team_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mask_istft_kernel -s 8 -c 1 -o gpurun_out/r1c_team1024_prof $CMD > gpurun_out/team_ncu.log 2>&1
tail -3 gpurun_out/team_ncu.log
