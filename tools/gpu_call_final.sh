#!/bin/bash
# Final validation call (1 GPU): smoke, the whole GPU suite, the default bench line with its cpu_baseline, DRAM traffic of every prefix launch
# of a pass with the final sources, --set full of the final fused prefix kernel, cfg4.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rA > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
grep -v "^PASSED" gpurun_out/r2j_pytest.log | grep -i "FAILED\|passed\|failed\|Error" | tail -8
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.log 2> gpurun_out/r2j_bench.err
cut -c1-200 gpurun_out/r2j_bench.log; tail -3 gpurun_out/r2j_bench.err
timeout 200 python tools/bench_config.py --cfg 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2j_bench_cfg4.log 2> gpurun_out/r2j_bench_cfg4.err
cut -c1-200 gpurun_out/r2j_bench_cfg4.log
C="bench.py --steps 1 --warmup 1 --no-cpu-baseline"
python $C > gpurun_out/r2j_plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:prefix_lazy --csv --log-file gpurun_out/r2j_prefix_traffic.csv python $C > gpurun_out/r2j_ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2j_lazy_fill python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2 > gpurun_out/r2j_ncu_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2j_lazy_tail python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120 > gpurun_out/r2j_ncu_tail.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"beam_combine|beam_candidates|ctc_log_softmax" -s 12 -c 3 -o gpurun_out/r2j_beam_kernels python $C > gpurun_out/r2j_ncu_beam.log 2>&1
ls -la gpurun_out | grep r2j_ | awk '{print $5, $9}'
