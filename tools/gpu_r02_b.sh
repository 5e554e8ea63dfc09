#!/bin/bash
# round 2, GPU call B: the two-stream attention forward (attention_fwd2.cu) - numerics, then timing against the first kernel
mkdir -p gpurun_out && cd /root/repo
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout=120 -p no:cacheprovider -x) > gpurun_out/b_attn_tests.log 2>&1
echo "attention kernel tests rc=$?"; tail -15 gpurun_out/b_attn_tests.log
for v in 1 2; do
  for shape in "8 64 256 16 64 1" "1 64 256 16 64 7" "4 64 128 8 64 1" "2 64 64 8 32 1" "4 32 1024 12 64 1"; do
    TA_ATTN_FWD=$v timeout 120 python tools/attn_fwd_only.py $shape 10 2>&1 | tail -1 | sed "s/^/v$v /"
  done
done | tee gpurun_out/b_attn_fwd_times.txt
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_engine_gpu.py tests/test_model_gpu.py -q --timeout=600 -p no:cacheprovider) > gpurun_out/b_engine_tests.log 2>&1
echo "engine tests rc=$?"; tail -8 gpurun_out/b_engine_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_c4.json 2> gpurun_out/b_bench_c4.err; tail -c 300 gpurun_out/b_bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'])
print({k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
