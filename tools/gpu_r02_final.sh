#!/bin/bash
# last call of the round: smoke(), the full GPU suite and the c4 line on the final build, all under tight timeouts
mkdir -p gpurun_out && cd /root/repo
timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fin_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 gpurun_out/fin_smoke.txt; exit 1; }
tail -1 gpurun_out/fin_smoke.txt
(timeout 420 python -m pytest tests -q -x -m gpu --timeout=120 -p no:cacheprovider) > gpurun_out/fin_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/fin_gpu_tests.log; grep -n "^E " gpurun_out/fin_gpu_tests.log | head -5 | cut -c1-300
TRADEAID_BENCH_DETAIL=gpurun_out/fin_gemm_detail_c4.txt timeout 240 python bench.py --no-cpu-baseline > gpurun_out/fin_bench_c4.json 2> gpurun_out/fin_bench_c4.err; echo "bench rc=$?"; tail -c 300 gpurun_out/fin_bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/fin_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'], d['e2e']['value'])
print({k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
