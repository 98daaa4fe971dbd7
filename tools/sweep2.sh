#!/bin/bash
for st in 0 500 1000 2000 3000 4000 6000; do
  GSS_STAGGER=$st python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('stagger=$st', 'synth_ms', round(r['ms_per_launch'],4), 'stft_ms', round(r['stft_kernel']['ms_per_launch'],4))
    elif 'rror' in l: print(l[:300])
"
done
