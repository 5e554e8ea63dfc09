#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
(timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_main_py_gpu.py -q --timeout=600 -p no:cacheprovider) > gpurun_out/n_tests.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/n_tests.log
timeout 400 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/n_bench_c5.json 2> gpurun_out/n_bench_c5.err; tail -c 400 gpurun_out/n_bench_c5.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n_bench_c5.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['generate'], d['clocks'])
PY
timeout 300 python bench.py --precision bf16x3 --batch 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_c4_bf16x3_b16.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/n_bench_c4_bf16x3_b16.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
