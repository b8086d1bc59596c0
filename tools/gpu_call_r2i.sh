#!/bin/bash
# GPU call (1 GPU): full suite with the wide-vocabulary beam kernels, Estrin vs Horner on latency-bound launches, default bench line,
# cfg3 with the new kernels, DRAM traffic of every prefix launch of a pass (final sources).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rA > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
grep -v "^PASSED" gpurun_out/r2i_pytest.log | grep -i "FAILED\|passed\|failed\|Error" | tail -8
for eb in 1000 0; do
  E2E_LAZY_ESTRIN_BELOW=$eb timeout 60 python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120
  E2E_LAZY_ESTRIN_BELOW=$eb timeout 60 python tools/bench_prefix.py --utts 6 --frames 825 --lazy 1 --poly 1 --plen 300
  E2E_LAZY_ESTRIN_BELOW=$eb timeout 60 python tools/bench_prefix.py --utts 600 --frames 600 --lazy 1 --poly 1 --plen 150
  E2E_LAZY_ESTRIN_BELOW=$eb timeout 60 python tools/bench_prefix.py --utts 256 --frames 875 --beam 16 --lazy 1 --poly 1 --plen 2
done > gpurun_out/r2i_prefix_micro_estrin.jsonl 2> gpurun_out/r2i_prefix_micro_estrin.err
cut -c1-290 gpurun_out/r2i_prefix_micro_estrin.jsonl; tail -2 gpurun_out/r2i_prefix_micro_estrin.err
timeout 60 python tools/bench_beam_kernels.py --utts 512 --vocab 10000 > gpurun_out/r2i_beam_micro_10k.jsonl 2>&1; cut -c1-700 gpurun_out/r2i_beam_micro_10k.jsonl
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
cut -c1-200 gpurun_out/r2i_bench.log; tail -3 gpurun_out/r2i_bench.err
E2E_LAZY_ESTRIN_BELOW=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_horner.log 2> gpurun_out/r2i_bench_horner.err
cut -c1-200 gpurun_out/r2i_bench_horner.log
timeout 300 python tools/bench_config.py --cfg 3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_bench_cfg3.log 2> gpurun_out/r2i_bench_cfg3.err
cut -c1-200 gpurun_out/r2i_bench_cfg3.log; tail -3 gpurun_out/r2i_bench_cfg3.err
C="bench.py --steps 1 --warmup 1 --no-cpu-baseline"
python $C > gpurun_out/r2i_plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:prefix_lazy --csv --log-file gpurun_out/r2i_prefix_traffic.csv python $C > gpurun_out/r2i_ncu_traffic.log 2>&1
tail -2 gpurun_out/r2i_ncu_traffic.log | cut -c1-200
