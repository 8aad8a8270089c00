#!/bin/bash
# Build a variant of the CUDA library with extra nvcc flags for A/B measurements:
#   scripts/build_variant.sh out.bin -DPNC_REG_MINB=11 ...      (load it with PNC_B200_LIB=out.bin)
set -e
out=$1; shift
cd "$(dirname "$0")/../pypnc_b200/csrc"
tmp=$(mktemp -d)
for f in capi dynamics ihwbc ihwbc_reg icprep managers; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o $tmp/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OLDPWD/$out" $tmp/*.o
rm -rf $tmp
