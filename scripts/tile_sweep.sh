#!/bin/bash
# Build timing variants of libOADG.so (dev container), then on the GPU box:
#   for v in gpurun_variants/*.so; do OADG_LIB=$PWD/$v python scripts/chain_stats.py 6 8 q | tail -2; done
# usage: scripts/tile_sweep.sh name1 "flags1" name2 "flags2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
rm -f gpurun_variants/*.so
build() {  # name, extra flags
  name=$1; shift
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
    -Xcompiler -ffp-contract=off -shared -fmad=false -I include -I oadg_b200/csrc $@ \
    oadg_b200/csrc/api.cu oadg_b200/csrc/saliency.cu oadg_b200/csrc/oamix.cu oadg_b200/csrc/oamix_sampler.cpp \
    oadg_b200/csrc/oaloss.cu oadg_b200/csrc/oaloss_tc.cu -o gpurun_variants/$name.so &
}
while [ $# -gt 1 ]; do build "$1" $2; shift; shift; done
wait
