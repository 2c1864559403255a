#!/bin/bash
# Build timing variants of libOADG.so with other tile sizes (dev container), then on the GPU box:
#   for v in gpurun_variants/*.so; do OADG_LIB=$PWD/$v python scripts/chain_stats.py 24 q | tail -1; done
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
build() {  # name, extra flags
  name=$1; shift
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
    -Xcompiler -ffp-contract=off -shared -fmad=false -I include -I oadg_b200/csrc "$@" \
    oadg_b200/csrc/api.cu oadg_b200/csrc/saliency.cu oadg_b200/csrc/oamix.cu oadg_b200/csrc/oamix_sampler.cpp \
    oadg_b200/csrc/oaloss.cu oadg_b200/csrc/oaloss_tc.cu -o gpurun_variants/$name.so
}
build base
build bbo128 -DOADG_BBO_TILE_W=128
build steppx128 -DOADG_STEP_TILE_W_PX=128
build catch256 -DOADG_CATCH_TILE_W=256
build bbo512 -DOADG_BBO_TILE_W=512
build tall -DOADG_BBO_TILE_H=32 -DOADG_STEP_TILE_H=32
