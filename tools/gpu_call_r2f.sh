#!/bin/bash
# GPU call (1 GPU): does oracle/_ref travel, full GPU suite, default bench line incl. the reference-based cpu_baseline, the reference arm on a
# small sample, micro-benchmarks of kernels (1) and (3), tuning-variant sweeps.
set -u
mkdir -p gpurun_out
ls -la oracle/_ref oracle/_ref/src > gpurun_out/r2f_ls.log 2>&1; cat gpurun_out/r2f_ls.log | head -12
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rA > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
grep -v "^PASSED" gpurun_out/r2f_pytest.log | grep -i "drop-in\|full-size\|long form\|V 10000\|FAILED\|passed\|failed\|rescored\|Error" | tail -20
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err
cut -c1-300 gpurun_out/r2f_bench.log; tail -3 gpurun_out/r2f_bench.err
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 --cpu-sample 8 --as-shipped-sample 2 > gpurun_out/r2f_bench_ref.log 2> gpurun_out/r2f_bench_ref.err
cut -c1-300 gpurun_out/r2f_bench_ref.log; tail -3 gpurun_out/r2f_bench_ref.err
for sh in "--utts 2620 --frames 180 --vocab 31" "--utts 2620 --frames 180 --vocab 31 --ragged 1" "--utts 512 --frames 180 --vocab 10000" "--utts 512 --frames 180 --vocab 10000 --ragged 1"; do
  timeout 60 python tools/bench_posterior.py $sh
done > gpurun_out/r2f_posterior_micro.jsonl 2> gpurun_out/r2f_posterior_micro.err
timeout 60 python tools/bench_beam_kernels.py >> gpurun_out/r2f_beam_micro.jsonl 2>> gpurun_out/r2f_posterior_micro.err
timeout 60 python tools/bench_beam_kernels.py --utts 512 --vocab 10000 >> gpurun_out/r2f_beam_micro.jsonl 2>> gpurun_out/r2f_posterior_micro.err
cut -c1-400 gpurun_out/r2f_posterior_micro.jsonl gpurun_out/r2f_beam_micro.jsonl; tail -3 gpurun_out/r2f_posterior_micro.err
timeout 200 python tools/sweep_prefix_variants.py attention > gpurun_out/r2f_sweep_attention.log 2>&1
timeout 200 python tools/sweep_prefix_variants.py lazy > gpurun_out/r2f_sweep_lazy.log 2>&1
cat gpurun_out/r2f_sweep_attention.log gpurun_out/r2f_sweep_lazy.log
for v in base ctx8 ctx16; do E2E_ASR_B200_LIB=e2e-asr-pytorch_b200/lib/variants/lib_$v.so timeout 60 python tools/bench_attention.py --utts 2620 --frames 180 --ragged 1 --kernels 1; E2E_ASR_B200_LIB=e2e-asr-pytorch_b200/lib/variants/lib_$v.so timeout 60 python tools/bench_attention.py --utts 600 --frames 824 --ragged 1 --kernels 1; done > gpurun_out/r2f_attention_ctx_variants.jsonl 2>&1
cut -c1-200 gpurun_out/r2f_attention_ctx_variants.jsonl
