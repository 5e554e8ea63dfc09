#!/bin/bash
# full GPU suite on the current build, bwd timing after the dropout templating, c2/c3 (PDL auto), c5 with clocks, c4 all-cross
mkdir -p gpurun_out && cd /root/repo
timeout 60 python tools/attn_bench.py 1 2 256 2 64 3 2 > gpurun_out/j_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/j_smoke.txt; exit 1; }
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7"; do timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done | tee gpurun_out/j_attn_times.txt
(time timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider) > gpurun_out/j_pytest.log 2>&1
tail -8 gpurun_out/j_pytest.log
timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_c2.json 2>/dev/null
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c3.json 2>/dev/null
timeout 400 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/j_bench_c5.json 2> gpurun_out/j_bench_c5.err
timeout 600 python bench.py --all-cross --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c4_allcross_b32.json 2> gpurun_out/j_bench_c4_allcross.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c4.json 2> gpurun_out/j_bench_c4.err
python - <<PY
import json
for f in ["j_bench_c2","j_bench_c3","j_bench_c5","j_bench_c4_allcross_b32","j_bench_c4"]:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), d.get('clocks'), round(d.get('hbm_bytes_allocated',0)/1e9,1))
        if 'breakdown_eager_step' in d: print("   ", {k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
        if 'generate' in d: print("   ", d['generate'], d['forward'])
    except Exception as e: print(f, "failed", e)
PY
