#!/bin/bash
# Tuning builds of the library for tools/sweep_prefix_variants.py (git-ignored; they travel to the GPU box with the snapshot).
#   bash tools/build_variants.sh && gpurun -- 'python tools/sweep_prefix_variants.py attention; python tools/sweep_prefix_variants.py lazy'
set -eu
cd "$(dirname "$0")/../e2e-asr-pytorch_b200/csrc"
rm -rf ../lib/variants; mkdir -p ../lib/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-extended-lambda -Xcompiler -fPIC -shared"
build() { name=$1; shift; nvcc $FLAGS "$@" -o ../lib/variants/lib_$name.so *.cu & }
build base
build ctx8  -DE2E_AF_CTX_FRAMES=8          # context product: 8 / 16 value rows in flight per thread (default 4)
build ctx16 -DE2E_AF_CTX_FRAMES=16
wait
build lazy_mb4 -DE2E_LAZY_MINBLOCKS=4      # fused prefix step: register cap 102 / 81 / 51 (default 6 CTAs of 160 threads: 68)
build lazy_mb5 -DE2E_LAZY_MINBLOCKS=5
build lazy_mb8 -DE2E_LAZY_MINBLOCKS=8
wait
ls -la ../lib/variants
