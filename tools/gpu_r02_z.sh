#!/bin/bash
# round-2 final artefacts: smoke(), full GPU suite, c4 / c2 / c3 / reference bench lines, HBM kernel table, ncu launch list of one
# eager step, ncu --set full of the attention backward kernels
mkdir -p gpurun_out && cd /root/repo
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 gpurun_out/z_smoke.txt; exit 1; }
tail -1 gpurun_out/z_smoke.txt
(timeout 1500 python -m pytest tests -q -x -m gpu --timeout=900 -p no:cacheprovider) > gpurun_out/z_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/z_gpu_tests.log
TRADEAID_BENCH_DETAIL=gpurun_out/z_gemm_detail_c4.txt timeout 900 python bench.py > gpurun_out/z_bench_c4.json 2> gpurun_out/z_bench_c4.err; echo "bench rc=$?"; tail -c 300 gpurun_out/z_bench_c4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/z_bench_c4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'], d['e2e']['value'], d['cpu_baseline'])
print({k:v['ms'] for k,v in d['breakdown_eager_step'].items()})
PY
timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --config c3 --no-cpu-baseline > gpurun_out/z_bench_c3.json 2> gpurun_out/z_bench_c3.err; echo "c3 rc=$?"
for c in c2 c3; do python -c "import json; d=json.loads(open('gpurun_out/z_bench_$c.json').read().strip().splitlines()[-1]); print('$c', d['ms_per_step'], d['value'])"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_bench_c4_reference.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/z_bench_c4_reference.json
timeout 600 python tools/hbm_kernels.py gpurun_out/z_hbm_kernels.json > gpurun_out/z_hbm_kernels.log 2>&1; echo "hbm rc=$?"; cut -c1-200 gpurun_out/z_hbm_kernels.log
L=$(python -c "import json; d=json.loads(open('gpurun_out/z_bench_c4.json').read().strip().splitlines()[-1]); print(d['gpu_launches']//d['steps'])")
timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/z_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((4 * L)) -c $L --csv --log-file gpurun_out/z_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/z_ncu_launches.log 2>&1
echo "ncu launches rc=$? (L=$L)"
for case in "self 8 64 256 16 64 1" "cross 1 64 256 16 64 7"; do set -- $case; n=$1; shift
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_bwd -s 4 -c 2 -f -o gpurun_out/z_bwd_$n python tools/attn_bwd_only.py $@ 3 > gpurun_out/z_ncu_bwd_$n.log 2>&1; tail -1 gpurun_out/z_ncu_bwd_$n.log; done
ls -la gpurun_out | grep " z_"; du -sh gpurun_out
