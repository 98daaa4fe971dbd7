#!/bin/bash
# tuning sweep (library built with -DGSS_TUNE): kernel times per variant
for sw in 4 9 10 11 12; do
  GSS_SYNTH_WARPS=$sw python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('synth_warps=$sw', 'synth_ms', round(r['ms_per_launch'],4), 'stft_ms', round(r['stft_kernel']['ms_per_launch'],4), 'step', round(d['ms_per_step'],4))
"
done
for sw in 12 14 16; do
  GSS_STFT_WARPS=$sw python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('stft_warps=$sw', 'synth_ms', round(r['ms_per_launch'],4), 'stft_ms', round(r['stft_kernel']['ms_per_launch'],4), 'step', round(d['ms_per_step'],4))
"
done
