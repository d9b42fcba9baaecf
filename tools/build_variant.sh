#!/bin/bash
# Tuning aid: builds the library with extra -D flags into build_variants/<name>/libfwi_b200.so (git-ignored).
# Tools under tools/ pick it up through FWI_VARIANT_LIB=<path>.   usage: tools/build_variant.sh np12 -DFD3_NP=12
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build_variants/$name
mkdir -p $out
for f in $root/full_waveform_inversion_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $out/$(basename $f .cu).o &
done
wait
nvcc -shared -o $out/libfwi_b200.so $out/*.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
rm -f $out/*.o
echo $out/libfwi_b200.so
