#!/bin/bash
# ncu --set full of the fused per-step prefix kernel on the micro-benchmark: a machine-filling launch and a tail-shaped one.
set -u
mkdir -p gpurun_out
A="tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2"
B="tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120"
timeout 300 python -m pytest tests/test_gpu_beam_kernels.py tests/test_gpu_decode.py -m gpu -q --timeout 180 -x > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv -lms 100 > gpurun_out/r2c_clocks.csv &
SMI=$!
python $A > gpurun_out/r2c_plain_a.log 2>&1 && python $B > gpurun_out/r2c_plain_b.log 2>&1 && python tools/bench_prefix.py --utts 64 --frames 825 --lazy 0 --poly 1 --plen 120 --skip-dead 1 > gpurun_out/r2c_plain_c.log 2>&1
kill $SMI
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2c_lazy_fill python $A > gpurun_out/r2c_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2c_lazy_tail python $B > gpurun_out/r2c_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_score -s 5 -c 1 -o gpurun_out/r2c_eager_tail python tools/bench_prefix.py --utts 64 --frames 825 --lazy 0 --poly 1 --plen 120 --skip-dead 1 > gpurun_out/r2c_ncu_c.log 2>&1
cat gpurun_out/r2c_plain_a.log gpurun_out/r2c_plain_b.log gpurun_out/r2c_plain_c.log | cut -c1-300
sort gpurun_out/r2c_clocks.csv | uniq -c | sort -rn | head -5
tail -2 gpurun_out/r2c_ncu_a.log gpurun_out/r2c_ncu_b.log gpurun_out/r2c_ncu_c.log
