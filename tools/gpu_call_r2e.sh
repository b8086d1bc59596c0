#!/bin/bash
# GPU call: full GPU suite (drop-in through the reference's caller, full-size golden with rescoring audit, wide posterior kernel),
# tile-size micro-benchmark, the default bench line with the reference-based cpu_baseline, the reference arm on a small sample,
# cfg3 / cfg4, and the ncu evidence (DRAM traffic of every prefix launch of a pass, launch list, --set full of the hand-written kernels).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rA > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
grep -v "^PASSED" gpurun_out/r2e_pytest.log | grep -i "drop-in\|full-size\|identical\|FAILED\|passed\|failed\|rescored" | tail -30
for t in 296 100000; do
  E2E_LAZY_SMALL_TILE_FROM=$t timeout 60 python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2
  E2E_LAZY_SMALL_TILE_FROM=$t timeout 60 python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 60
  E2E_LAZY_SMALL_TILE_FROM=$t timeout 60 python tools/bench_prefix.py --utts 1200 --frames 300 --lazy 1 --poly 1 --plen 60
  E2E_LAZY_SMALL_TILE_FROM=$t timeout 60 python tools/bench_prefix.py --utts 500 --frames 400 --lazy 1 --poly 1 --plen 100
done > gpurun_out/r2e_prefix_micro_tiles.jsonl 2> gpurun_out/r2e_prefix_micro_tiles.err
cut -c1-300 gpurun_out/r2e_prefix_micro_tiles.jsonl
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
cut -c1-400 gpurun_out/r2e_bench.log; tail -3 gpurun_out/r2e_bench.err
timeout 500 python bench.py --impl reference --steps 4 --warmup 1 --cpu-sample 16 --as-shipped-sample 2 > gpurun_out/r2e_bench_ref.log 2> gpurun_out/r2e_bench_ref.err
cut -c1-600 gpurun_out/r2e_bench_ref.log; tail -3 gpurun_out/r2e_bench_ref.err
timeout 200 python tools/bench_config.py --cfg 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2e_bench_cfg4.log 2> gpurun_out/r2e_bench_cfg4.err
timeout 300 python tools/bench_config.py --cfg 3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_cfg3.log 2> gpurun_out/r2e_bench_cfg3.err
cut -c1-300 gpurun_out/r2e_bench_cfg4.log gpurun_out/r2e_bench_cfg3.log; tail -3 gpurun_out/r2e_bench_cfg4.err gpurun_out/r2e_bench_cfg3.err
# ---- ncu (the same bench command has just exited 0 above)
C="bench.py --steps 1 --warmup 1 --no-cpu-baseline"
python $C > gpurun_out/r2e_plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:prefix_lazy --csv --log-file gpurun_out/r2e_prefix_traffic.csv python $C > gpurun_out/r2e_ncu_traffic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 18196 -c 18196 --csv --log-file gpurun_out/r2e_launches.csv python $C > gpurun_out/r2e_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"beam_combine|beam_candidates|attention_energy|attention_softmax|prefix_lazy" -s 40 -c 10 -o gpurun_out/r2e_step_kernels python $C > gpurun_out/r2e_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ctc_log_softmax|ctc_init" -c 2 -o gpurun_out/r2e_posterior python $C > gpurun_out/r2e_ncu_post.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2e_lazy_fill python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2 > gpurun_out/r2e_ncu_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefix_lazy -s 5 -c 1 -o gpurun_out/r2e_lazy_tail python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120 > gpurun_out/r2e_ncu_tail.log 2>&1
ls -la gpurun_out/ | grep r2e_ | awk '{print $5, $9}'
