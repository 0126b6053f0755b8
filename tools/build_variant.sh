#!/bin/bash
# Build an alternative libsmk.so for a tuning experiment (selected at run time with SMK_LIB=<path>):
#   tools/build_variant.sh mb5 "-DSMK_MIN_BLOCKS_FAST=5"
# -> simplemoc-kernel_b200/lib/variants/libsmk_mb5.so   (git-ignored; travels to the GPU box with gpurun)
set -e
NAME=$1; shift
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/simplemoc-kernel_b200/lib/variants
mkdir -p $OUT
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-ffp-contract=off,-Wall "$@" --shared $HERE/simplemoc-kernel_b200/csrc/smk_api.cu \
    -o $OUT/libsmk_$NAME.so
echo built $OUT/libsmk_$NAME.so
