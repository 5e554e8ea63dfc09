#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dataprep_gpu.py tests/test_zz_demo_prep_gpu.py -q --timeout=600 -p no:cacheprovider) > gpurun_out/k_tests.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/k_tests.log
tail -12 gpurun_out/parity_r02.jsonl | cut -c1-420
timeout 300 python bench.py --precision bf16x3 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_c4_bf16x3_b16.json 2> gpurun_out/k_bench_c4_bf16x3.err; tail -c 300 gpurun_out/k_bench_c4_bf16x3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/k_bench_c4_bf16x3_b16.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
