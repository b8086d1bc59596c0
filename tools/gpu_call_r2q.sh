#!/bin/bash
# GPU call: attention energy kernel, reciprocal shared by two channels of one hypothesis (variant libraries) vs the default.
set -u
mkdir -p gpurun_out
V=e2e-asr-pytorch_b200/lib/variants/lib_pairrcp.so
V2=e2e-asr-pytorch_b200/lib/variants/lib_pairrcp_mb2.so
E2E_ASR_B200_LIB=$V timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py tests/test_gpu_fullsize_golden.py tests/test_gpu_dropin_reference.py -m gpu -q --timeout 400 > gpurun_out/r2q_pytest_pair.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest_pair.log; tail -4 gpurun_out/r2q_pytest_pair.log
for lib in "" $V $V2; do
  for shape in "--utts 2620 --frames 180 --ragged 1" "--utts 600 --frames 824 --ragged 1" "--utts 100 --frames 824 --ragged 1"; do
    E2E_ASR_B200_LIB=$lib timeout 120 python tools/bench_attention.py $shape --kernels 1
  done
done > gpurun_out/r2q_attention_micro.jsonl 2> gpurun_out/r2q_attention_micro.err
cut -c1-120 gpurun_out/r2q_attention_micro.jsonl; tail -2 gpurun_out/r2q_attention_micro.err
E2E_ASR_B200_LIB=$V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_pair.log 2> gpurun_out/r2q_bench_pair.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_base.log 2> gpurun_out/r2q_bench_base.err
python - <<'PY'
import json
for n in ('pair','base'):
    try:
        d=json.loads(open('gpurun_out/r2q_bench_%s.log'%n).read().strip().splitlines()[-1])
        print(n,'value',d['value'],'e2e',d['e2e']['value'],d['nbest_parity'],d.get('phases_ms'))
    except Exception as e:
        print(n,'failed',e)
PY
tail -3 gpurun_out/r2q_bench_pair.err
