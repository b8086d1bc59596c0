#!/bin/bash
# GPU call: fused per-step prefix kernel v2 (helper warps): parity, micro-benchmark, decode with table / polynomial math; beam-kernel tests.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 240 -rA > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
for poly in 0 1; do
  timeout 60 python tools/bench_prefix.py --utts 2620 --lazy 1 --poly $poly --plen 2
  timeout 60 python tools/bench_prefix.py --utts 2620 --plen 60 --lazy 1 --poly $poly
  timeout 60 python tools/bench_prefix.py --utts 1200 --frames 300 --plen 60 --lazy 1 --poly $poly
  timeout 60 python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly $poly --plen 120
  timeout 60 python tools/bench_prefix.py --utts 256 --frames 875 --beam 16 --lazy 1 --poly $poly --plen 2
done > gpurun_out/r2d_prefix_micro.jsonl 2> gpurun_out/r2d_prefix_micro.err
timeout 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_lazy_lut.log 2> gpurun_out/r2d_bench_lazy_lut.err
timeout 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefix-math poly > gpurun_out/r2d_bench_lazy_poly.log 2> gpurun_out/r2d_bench_lazy_poly.err
E2E_LAZY_SMALL_TILE_FROM=100000 timeout 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefix-math poly > gpurun_out/r2d_bench_lazy_poly_t32.log 2> gpurun_out/r2d_bench_lazy_poly_t32.err
grep -v "^PASSED" gpurun_out/r2d_pytest.log | tail -25
cut -c1-330 gpurun_out/r2d_prefix_micro.jsonl
tail -3 gpurun_out/r2d_prefix_micro.err
for f in lut poly poly_t32; do cut -c1-200 gpurun_out/r2d_bench_lazy_$f.log; tail -3 gpurun_out/r2d_bench_lazy_$f.err; done
