#!/bin/bash
# usage: tools/build_variant2.sh <name> <source file in csrc, e.g. wpf1920.cu> "<-D flags>"   -> _ab/<name>.so (that file recompiled with the flags, other objects reused)
set -e
cd "$(dirname "$0")/.."
mkdir -p _ab
B=mlx_swift_audio_b200/_build
stem="${2%.*}"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v $3 \
  -x cu -c mlx_swift_audio_b200/csrc/$2 -o _ab/${stem}_$1.o 2>&1 | grep -E "registers|spill" | sort | uniq -c
objs=""
for o in host_tables frontend wpf1920 vocoder tc_frontend generic_stft capi; do
  if [ "$o" = "$stem" ]; then objs="$objs _ab/${stem}_$1.o"; else objs="$objs $B/$o.o"; fi
done
nvcc -shared -o _ab/$1.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static
echo built _ab/$1.so
