#!/bin/bash
# Final validation call of round 2 (1 GPU): smoke, the whole GPU suite, the default bench line with its cpu_baseline, DRAM traffic of every
# prefix launch of a pass with the final sources, --set full of the final fused prefix kernel (fill / tail shapes), instruction micro-benchmark.
set -u
mkdir -p gpurun_out
build/lat > gpurun_out/r2r_lat.txt 2>&1; tail -4 gpurun_out/r2r_lat.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; tail -2 gpurun_out/r2r_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rA > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
grep -v "^PASSED" gpurun_out/r2r_pytest.log | grep -i "FAILED\|passed\|failed\|Error" | tail -8
C="bench.py --steps 1 --warmup 1 --no-cpu-baseline"
python $C > gpurun_out/r2r_plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:prefix_lazy --csv --log-file gpurun_out/r2r_prefix_traffic.csv python $C > gpurun_out/r2r_ncu_traffic.log 2>&1
tail -1 gpurun_out/r2r_ncu_traffic.log | cut -c1-200
python tools/summarize_ncu.py traffic gpurun_out/r2r_prefix_traffic.csv --first 660 --count 660 > gpurun_out/r2r_traffic_summary.json 2> gpurun_out/r2r_traffic_summary.err; cut -c1-300 gpurun_out/r2r_traffic_summary.json; tail -2 gpurun_out/r2r_traffic_summary.err
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench.log 2> gpurun_out/r2r_bench.err
cut -c1-200 gpurun_out/r2r_bench.log; tail -3 gpurun_out/r2r_bench.err
timeout 60 python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2 > gpurun_out/r2r_plain_fill.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2r_lazy_fill python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2 > gpurun_out/r2r_ncu_fill.log 2>&1
timeout 60 python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120 > gpurun_out/r2r_plain_tail.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2r_lazy_tail python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120 > gpurun_out/r2r_ncu_tail.log 2>&1
ls -la gpurun_out | grep r2r_ | awk '{print $5, $9}'
