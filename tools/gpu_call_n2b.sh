#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_n2.log 2> gpurun_out/r2m_bench_n2.err
cut -c1-200 gpurun_out/r2m_bench_n2.log; tail -4 gpurun_out/r2m_bench_n2.err | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_shard_pack.py -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
