#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
TRADEAID_BENCH_DETAIL=gpurun_out/g_gemm_detail_c4.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_c4.json 2> gpurun_out/g_bench_c4.err; tail -c 300 gpurun_out/g_bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/g_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'])
print({k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7"; do timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done | tee gpurun_out/g_attn_times.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/g_bwd_self python tools/attn_bwd_only.py 8 64 256 16 64 1 3 > gpurun_out/g_ncu_bwd_self.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/g_bwd_cross python tools/attn_bwd_only.py 1 64 256 16 64 7 3 > gpurun_out/g_ncu_bwd_cross.log 2>&1
tail -2 gpurun_out/g_ncu_bwd_cross.log
