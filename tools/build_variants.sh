#!/bin/bash
# Tuning builds of librtclj_b200 with other compile-time parameters.
# usage: tools/build_variants.sh NAME:"-DFOO=1 -DBAR=2" ...
# Output: raytracing-clj_b200/csrc/build/variants/librtclj_NAME.so  (use with RTCLJ_LIB=...)
set -e
cd "$(dirname "$0")/../raytracing-clj_b200/csrc"
mkdir -p build/variants
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -c -o build/variants/host.o rtclj_host.cpp
for v in "$@"; do
  NAME=${v%%:*}; DEFS=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC \
       $DEFS -c -o build/variants/abi_$NAME.o rtclj_abi.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librtclj_$NAME.so \
       build/variants/abi_$NAME.o build/variants/host.o -cudart static
  echo "built $NAME ($DEFS)"
done
