#!/bin/bash
# Build the library with extra nvcc flags into _ab_<name>/libpafuse_b200.so (git-ignored, travels to the GPU box) for A/B
# measurements with PAFUSE_LIB=...:   tools/build_variant.sh epi8 -DPAFUSE_EPI_WARPS=8 ; tools/build_variant.sh ablate -DPAFUSE_ABLATE
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=_ab_$name
mkdir -p $out
flags="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
for s in pafuse_api gemm_tcgen05 attention attention_tc elementwise; do
  nvcc $flags "$@" -c pafuse_b200/csrc/$s.cu -o $out/$s.o &
done
wait
nvcc -shared -o $out/libpafuse_b200.so $out/*.o -cudart static
rm -f $out/*.o
echo $out/libpafuse_b200.so
