#!/bin/bash
# usage: tools/build_variant.sh <name> "<-D flags>"   -> _ab/<name>.so (frontend.cu recompiled with the flags, other objects reused)
set -e
cd "$(dirname "$0")/.."
mkdir -p _ab
B=mlx_swift_audio_b200/_build
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr $2 \
  -x cu -c mlx_swift_audio_b200/csrc/frontend.cu -o _ab/frontend_$1.o
nvcc -shared -o _ab/$1.so $B/host_tables.o _ab/frontend_$1.o $B/vocoder.o $B/tc_frontend.o $B/generic_stft.o $B/capi.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built _ab/$1.so
