#!/bin/bash
# compile libgss.so and print registers / spills / SASS size per kernel (filter with $1)
cd "$(dirname "$0")/../gan_sass_tf_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -shared -Xcompiler -fPIC -o /tmp/libgss_regs.so csrc/gss_api.cu 2>&1 \
 | grep -E "error|Compiling entry|Used|spill" | sed -E 's/ptxas info\s+: //g' \
 | awk '/error/{print} /Compiling/{name=$4} /spill/{sp=$5" "$9} /Used/{print name, $2, "regs; spill st/ld", sp}' | grep -E "${1:-.}" | cut -c1-160
cuobjdump -sass /tmp/libgss_regs.so | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ /{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -rn | grep -E "${1:-.}" | head -8
