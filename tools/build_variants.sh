#!/bin/bash
# Tuning builds of the wavefront kernel: librtclj_b200 with other (threads, slots) per CTA.
# Output: raytracing-clj_b200/csrc/build/variants/librtclj_T<threads>_S<slots>.so  (use with RTCLJ_LIB=...)
set -e
cd "$(dirname "$0")/../raytracing-clj_b200/csrc"
mkdir -p build/variants
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c -o build/variants/host.o rtclj_host.cpp
for v in "$@"; do
  T=${v%%:*}; S=${v##*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC \
       -DRTCLJ_WAVE_THREADS=$T -DRTCLJ_WAVE_SLOTS=$S -c -o build/variants/abi_T${T}_S${S}.o rtclj_abi.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librtclj_T${T}_S${S}.so \
       build/variants/abi_T${T}_S${S}.o build/variants/host.o -cudart static
  echo built T=$T S=$S
done
