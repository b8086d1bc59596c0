#!/bin/bash
# 2-GPU call: tests of the re-written beam kernels / host copy path, bench at N=1 and N=2 on the same box.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_beam_kernels.py tests/test_gpu_decode.py tests/test_gpu_host_features.py tests/test_gpu_kernels.py tests/test_gpu_fullsize_golden.py -m gpu -q --timeout 400 -x > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -4 gpurun_out/r2g_pytest.log
timeout 60 python tools/bench_beam_kernels.py > gpurun_out/r2g_beam_micro.jsonl 2>&1; cut -c1-500 gpurun_out/r2g_beam_micro.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_n1.log 2> gpurun_out/r2g_bench_n1.err
cut -c1-200 gpurun_out/r2g_bench_n1.log; tail -3 gpurun_out/r2g_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_n2.log 2> gpurun_out/r2g_bench_n2.err
cut -c1-200 gpurun_out/r2g_bench_n2.log; tail -5 gpurun_out/r2g_bench_n2.err
