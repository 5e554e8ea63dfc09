#!/bin/bash
# GPU tests + c4 bench line with per-shape GEMM detail. usage: tools/gpu_quick.sh <tag> [pytest args]
tag=${1:-rXX}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu "$@" > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_gpu_tests.log
TRADEAID_BENCH_DETAIL=gpurun_out/${tag}_gemm_detail_c4.txt timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline']['achieved'], d['clocks'])
for k,v in d['breakdown_eager_step'].items(): print(k, v)
PY
head -24 gpurun_out/${tag}_gemm_detail_c4.txt
