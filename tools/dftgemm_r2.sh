#!/bin/bash
# BASELINE config C5: radix FFT (libgss) vs the tensor-core DFT-as-GEMM form (cuBLAS stand-in), B = 1024 x 3 s, hop N/4,
# analysis and synthesis side, every FFT size of the sweep.   gpurun --timeout 900 -- 'bash tools/dftgemm_r2.sh'
for N in 256 512 1024 2048 4096; do
  echo "== C5: N=$N H=$((N/4)) B=1024 n=48000"
  python tools/kbench.py $N $((N/4)) 1024 48000 --dftgemm --torch 2>&1 | grep -E "^stft  |^istft  |GEMM|cuFFT|libgss stft"
done
