#!/bin/bash
# GPU call: software-pipelined chains + 64-frame tiles for the deep tail: parity (kernel chain tests + decode suites), micro-benchmark, bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prefix_poly.py tests/test_gpu_decode.py tests/test_gpu_fullsize_golden.py -m gpu -q --timeout 400 -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -4 gpurun_out/r2o_pytest.log
for bt in 200 0; do
  E2E_LAZY_BIG_TILE_BELOW=$bt timeout 60 python tools/bench_prefix.py --utts 64 --frames 825 --lazy 1 --poly 1 --plen 120
  E2E_LAZY_BIG_TILE_BELOW=$bt timeout 60 python tools/bench_prefix.py --utts 6 --frames 825 --lazy 1 --poly 1 --plen 300
done > gpurun_out/r2o_prefix_micro.jsonl 2> gpurun_out/r2o_prefix_micro.err
timeout 60 python tools/bench_prefix.py --utts 2620 --lazy 1 --poly 1 --plen 2 >> gpurun_out/r2o_prefix_micro.jsonl 2>> gpurun_out/r2o_prefix_micro.err
timeout 60 python tools/bench_prefix.py --utts 600 --frames 600 --lazy 1 --poly 1 --plen 150 >> gpurun_out/r2o_prefix_micro.jsonl 2>> gpurun_out/r2o_prefix_micro.err
cut -c1-290 gpurun_out/r2o_prefix_micro.jsonl; tail -2 gpurun_out/r2o_prefix_micro.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench.log 2> gpurun_out/r2o_bench.err
cut -c1-200 gpurun_out/r2o_bench.log; tail -3 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.log').read().strip().splitlines()[-1]); r=d['roofline']
print('frac',r['frac'],'kms',r['kernel_ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],d['nbest_parity'])
PY
