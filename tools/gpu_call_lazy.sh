#!/bin/bash
# GPU call: parity of the fused per-step prefix kernel (lazy state evaluation), its micro-benchmark and the decode with it.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -rA -x > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
for poly in 0 1 2; do
  for lazy in 0 1; do
    timeout 60 python tools/bench_prefix.py --utts 2620 --lazy $lazy --poly $poly --plen 2
    timeout 60 python tools/bench_prefix.py --utts 2620 --plen 60 --skip-dead 1 --lazy $lazy --poly $poly
    timeout 60 python tools/bench_prefix.py --utts 64 --frames 825 --lazy $lazy --poly $poly --plen 2
    timeout 60 python tools/bench_prefix.py --utts 6 --frames 825 --lazy $lazy --poly $poly --plen 2
    timeout 60 python tools/bench_prefix.py --utts 256 --frames 875 --beam 16 --lazy $lazy --poly $poly --plen 2
  done
done > gpurun_out/r2b_prefix_micro.jsonl 2> gpurun_out/r2b_prefix_micro.err
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_lazy_lut.log 2> gpurun_out/r2b_bench_lazy_lut.err
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefix-math poly > gpurun_out/r2b_bench_lazy_poly.log 2> gpurun_out/r2b_bench_lazy_poly.err
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefix-math poly_estrin > gpurun_out/r2b_bench_lazy_estrin.log 2> gpurun_out/r2b_bench_lazy_estrin.err
tail -5 gpurun_out/r2b_pytest.log
cut -c1-330 gpurun_out/r2b_prefix_micro.jsonl
tail -3 gpurun_out/r2b_prefix_micro.err
for f in lut poly estrin; do cut -c1-200 gpurun_out/r2b_bench_lazy_$f.log; tail -2 gpurun_out/r2b_bench_lazy_$f.err; done
