#!/bin/bash
# GPU call: kernel (1) in its final forms (narrow rows, wide rows) and the warp-per-utterance initial-state kernel: timing + ncu --set full.
set -u
mkdir -p gpurun_out
for shape in "--utts 2620 --frames 180 --vocab 31 --ragged 1" "--utts 512 --frames 180 --vocab 10000 --ragged 1"; do
  timeout 100 python tools/bench_posterior.py $shape
done > gpurun_out/r2t_posterior_micro.jsonl 2> gpurun_out/r2t_posterior_micro.err
cut -c1-250 gpurun_out/r2t_posterior_micro.jsonl; tail -2 gpurun_out/r2t_posterior_micro.err
ncu --set full --clock-control none --import-source on -k regex:"ctc_log_softmax|ctc_init_state" -s 3 -c 2 -o gpurun_out/r2t_posterior_narrow python tools/bench_posterior.py --utts 2620 --frames 180 --vocab 31 --ragged 1 --steps 2 > gpurun_out/r2t_ncu_narrow.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ctc_log_softmax" -s 3 -c 1 -o gpurun_out/r2t_posterior_wide python tools/bench_posterior.py --utts 512 --frames 180 --vocab 10000 --ragged 1 --steps 2 > gpurun_out/r2t_ncu_wide.log 2>&1
tail -2 gpurun_out/r2t_ncu_narrow.log gpurun_out/r2t_ncu_wide.log | cut -c1-200
ls -la gpurun_out | grep r2t_ | awk '{print $5, $9}'
