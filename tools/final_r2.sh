#!/bin/bash
# Round-end evidence run (one GPU): tests, bench (both arms), smoke, per-shape kernel timings, ncu launch list + full capture.
#   gpurun --timeout 1500 -- 'bash tools/final_r2.sh r2g'
set -u
TAG=${1:-r2g}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; tail -n 2 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -n 1 $O/${TAG}_smoke.log
{
  echo "== C2: N=512 H=128 B=256 n=48000";   python tools/kbench.py 512 128 256 48000 --torch
  echo "== C4 share at 8 GPUs: N=512 H=128 B=1024 n=64000"; python tools/kbench.py 512 128 1024 64000
  for N in 256 512 1024 2048 4096; do echo "== C5: N=$N H=$((N/4)) B=1024 n=48000"; python tools/kbench.py $N $((N/4)) 1024 48000; done
  echo "== reference default: N=256 H=128 B=1024 n=48000"; python tools/kbench.py 256 128 1024 48000
  echo "== C3: N=1024 H=256 B=1 n=960000"; python tools/kbench.py 1024 256 1 960000
} 2>&1 | grep -v Warn > $O/${TAG}_kbench.txt
bash tools/profile_r2.sh $TAG
