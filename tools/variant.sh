#!/bin/bash
# tuning builds: tools/variant.sh <tag> [extra nvcc -D flags ...]  ->  gan_sass_tf_b200/lib/libgss_<tag>.so
# Only part 1 (the N = 512 streaming kernels, restricted by GSS_QUICK to the C2 instances) is recompiled;
# the other parts are cached under /tmp/gobj.  Select the result with GSS_LIB=... (see _native.py).
set -e
cd "$(dirname "$0")/../gan_sass_tf_b200/csrc"
tag=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fno-gnu-unique"
mkdir -p /tmp/gobj
for k in 0 2 3 4; do [ -f /tmp/gobj/p$k.o ] || nvcc $F -DGSS_PART=$k -c gss_api.cu -o /tmp/gobj/p$k.o 2>/dev/null & done
nvcc $F -DGSS_PART=1 -DGSS_QUICK -DGSS_TUNE "$@" -Xptxas -v -c gss_api.cu -o /tmp/gobj/p1_$tag.o 2>&1 | grep -E "error|Compiling entry|Used|spill" | sed -E 's/ptxas info\s+: //g' \
 | awk '/error/{print} /Compiling/{name=$4} /spill/{sp=$5" "$9} /Used/{print substr(name,1,60), $2, "regs; spill st/ld", sp}'
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libgss_$tag.so /tmp/gobj/p0.o /tmp/gobj/p1_$tag.o /tmp/gobj/p2.o /tmp/gobj/p3.o /tmp/gobj/p4.o
echo built lib/libgss_$tag.so
