#!/bin/bash
# tuning builds of the team kernels: tools/variant4.sh <tag> [extra nvcc -D flags ...] -> lib/libgss_<tag>.so
# Only part 3 (team kernels N >= 2048) is recompiled; the other parts are cached under /tmp/gobj (see tools/variant.sh).
set -e
cd "$(dirname "$0")/../gan_sass_tf_b200/csrc"
tag=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fno-gnu-unique"
mkdir -p /tmp/gobj
for k in 0 1 2 3; do [ -f /tmp/gobj/p$k.o ] || nvcc $F -DGSS_PART=$k -c gss_api.cu -o /tmp/gobj/p$k.o 2>/dev/null & done
nvcc $F -DGSS_PART=4 "$@" -Xptxas -v -c gss_api.cu -o /tmp/gobj/p4_$tag.o 2>&1 | grep -E "error|Compiling entry|Used|spill" | sed -E 's/ptxas info\s+: //g' \
 | awk '/error/{print} /Compiling/{name=$4} /spill/{sp=$5" "$9} /Used/{print substr(name,1,75), $2, "regs; spill st/ld", sp}' | grep -E "error|feat_kernelILi(2048|4096)ELi4ELi1"
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libgss_$tag.so /tmp/gobj/p0.o /tmp/gobj/p1.o /tmp/gobj/p2.o /tmp/gobj/p3.o /tmp/gobj/p4_$tag.o
echo built lib/libgss_$tag.so
