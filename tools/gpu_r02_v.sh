#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
timeout 90 python tools/attn_bench.py 1 8 256 16 64 2 2 > gpurun_out/v_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -5 gpurun_out/v_smoke.txt; exit 1; }
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7"; do timeout 120 python tools/attn_bwd_prof.py $shape 2>&1 | tail -4; done | tee gpurun_out/v_bwd_prof.txt
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -x -k "attention or attn" --timeout=300 -p no:cacheprovider) > gpurun_out/v_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/v_gpu_tests.log
TRADEAID_BENCH_DETAIL=gpurun_out/v_gemm_detail_c4.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_c4.json 2> gpurun_out/v_bench_c4.err; tail -c 300 gpurun_out/v_bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/v_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'])
print({k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
