#!/bin/bash
# 8-GPU call: the weak-scaling bench line at N=8 (8 x 2620 utterances) with the per-phase breakdown and the N-GPU == 1-GPU check.
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_n8.log 2> gpurun_out/r2s_bench_n8.err
cut -c1-200 gpurun_out/r2s_bench_n8.log; tail -5 gpurun_out/r2s_bench_n8.err
nproc; grep -c processor /proc/cpuinfo
