#!/bin/bash
# GPU call: CUDA-graph replay of the LSTM stacks — parity (end-to-end suites), host vs device time of the tail, bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fullsize_golden.py tests/test_gpu_dropin_reference.py tests/test_gpu_lm_f16x2.py tests/test_gpu_host_features.py -m gpu -q --timeout 400 -x > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -6 gpurun_out/r2l_pytest.log
timeout 200 python tools/profile_tail.py --from-step 300 > gpurun_out/r2l_profile_tail.json 2> gpurun_out/r2l_profile_tail.err; tail -3 gpurun_out/r2l_profile_tail.err; head -14 gpurun_out/r2l_profile_tail.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.log 2> gpurun_out/r2l_bench.err
cut -c1-200 gpurun_out/r2l_bench.log; tail -3 gpurun_out/r2l_bench.err
E2E_STEP_GRAPHS=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_nograph.log 2> gpurun_out/r2l_bench_nograph.err
cut -c1-200 gpurun_out/r2l_bench_nograph.log
