#!/bin/bash
# A/B builds: recompile ONE translation unit of libptg_b200 with extra -D flags and link it with the stock objects.
#   profiles/diag/build_variant.sh <name> <source.cu> "<-DFOO=1 -DBAR=2>"
# -> postpic_b200/lib/variants/libptg_<name>.so   (use with PTG_B200_LIB=... ; git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/../.."
name=$1; src=$2; defs=$3
mkdir -p postpic_b200/lib/variants
obj=postpic_b200/lib/variants/${name}_$(basename $src .cu).o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2 \
     --expt-relaxed-constexpr $defs -I include -c postpic_b200/csrc/$src -o $obj
objs=""
for o in postpic_b200/lib/obj/*.o; do
  if [ "$(basename $o)" != "$(basename $src .cu).o" ]; then objs="$objs $o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o postpic_b200/lib/variants/libptg_${name}.so $obj $objs
rm -f $obj
echo postpic_b200/lib/variants/libptg_${name}.so
