#!/bin/bash
# First GPU call of the next round: everything round 1 prepared without GPU minutes, in one go (~10 min of box time).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/r2_*.log; nothing here changes a default.
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite including the tests that are still waiting for their first run (poly log-add-exp)
E2E_UNVALIDATED_TESTS=1 timeout 400 python -m pytest tests -m gpu -q --timeout 120 -rA > gpurun_out/r2_pytest_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_all.log
# 2. prefix-score kernel, stand-alone: table vs polynomial log-add-exp on the three launch shapes of DESIGN.md §6
for poly in 0 1 2; do
  timeout 40 python tools/bench_prefix.py --utts 2620 --poly $poly
  timeout 40 python tools/bench_prefix.py --utts 2620 --plen 60 --skip-dead 1 --poly $poly
  timeout 40 python tools/bench_prefix.py --utts 64 --frames 825 --poly $poly
  timeout 40 python tools/bench_prefix.py --utts 256 --frames 875 --beam 16 --poly $poly
done > gpurun_out/r2_prefix_math_micro.jsonl 2> gpurun_out/r2_prefix_math_micro.err
# 2b. attention context product: direct (default) vs bulk-copy staged value tiles; parity of the staged variant first
E2E_AF_CTX_STAGED=1 timeout 120 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q -k "attention or decode_batch_matches or order_independent" > gpurun_out/r2_pytest_ctx_staged.log 2>&1
for staged in 0 1; do
  E2E_AF_CTX_STAGED=$staged timeout 40 python tools/bench_attention.py --utts 2620 --frames 180 --ragged 1 --kernels 1
  E2E_AF_CTX_STAGED=$staged timeout 40 python tools/bench_attention.py --utts 600 --frames 824 --ragged 1 --kernels 1
done > gpurun_out/r2_attention_ctx_staged.jsonl 2> gpurun_out/r2_attention_ctx_staged.err
E2E_AF_CTX_STAGED=1 timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ctx_staged.log 2> gpurun_out/r2_bench_ctx_staged.err
# 3. the decode with the polynomial evaluator (roofline.frac in-decode) next to the default
timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_default.log 2> gpurun_out/r2_bench_default.err
timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefix-math poly > gpurun_out/r2_bench_poly.log 2> gpurun_out/r2_bench_poly.err
timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ragged-h2d --ragged-gather > gpurun_out/r2_bench_ragged_h2d.log 2> gpurun_out/r2_bench_ragged_h2d.err
# 4. the two BASELINE configurations the bench does not run (one pass each; cfg3 in memory-budgeted batches)
timeout 150 python tools/bench_config.py --cfg 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_cfg4.log 2> gpurun_out/r2_bench_cfg4.err
timeout 240 python tools/bench_config.py --cfg 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.log 2> gpurun_out/r2_bench_cfg3.err
tail -3 gpurun_out/r2_pytest_all.log; tail -2 gpurun_out/r2_pytest_ctx_staged.log; cut -c1-200 gpurun_out/r2_attention_ctx_staged.jsonl
cat gpurun_out/r2_prefix_math_micro.jsonl | cut -c1-260
cut -c1-200 gpurun_out/r2_bench_default.log gpurun_out/r2_bench_poly.log gpurun_out/r2_bench_ragged_h2d.log gpurun_out/r2_bench_cfg4.log gpurun_out/r2_bench_cfg3.log
