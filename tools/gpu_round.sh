#!/bin/bash
# Round artefacts in one gpurun call: GPU tests, c4 bench line (+cpu_baseline), reference arm, HBM kernel table, ncu launch
# list of one eager step, ncu --set full of the HBM kernels.   usage: tools/gpu_round.sh <tag>   (outputs: gpurun_out/<tag>_*)
tag=${1:-rXX}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_gpu_tests.log; fi
TRADEAID_BENCH_DETAIL=gpurun_out/${tag}_gemm_detail_c4.txt timeout 900 python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/${tag}_bench_c4.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c4_reference.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_c4_reference.json
timeout 600 python bench.py --config c5 --steps 3 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; echo "c5 rc=$?"
timeout 600 python tools/hbm_kernels.py gpurun_out/${tag}_hbm_kernels.json > gpurun_out/${tag}_hbm_kernels.log 2>&1; echo "hbm rc=$?"; cat gpurun_out/${tag}_hbm_kernels.log | cut -c1-220
# ncu: launch list of one eager step (7 eager steps run: skip 4, capture one), then one --set full launch of every HBM kernel
L=$(python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_c4.json').read().strip().splitlines()[-1]); print(d['gpu_launches']//d['steps'])")
timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((4 * L)) -c $L --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu launches rc=$? (L=$L)"
# the .ncu-rep stays on the box (gpurun_out/ is capped at 64 MiB): only its csv pages and the one-screen summary come back
if [ -z "$SKIP_HBM_NCU" ]; then
HBM_ONCE=1 timeout 300 python tools/hbm_kernels.py > /dev/null 2>&1 && \
HBM_ONCE=1 timeout 900 ncu --set full --clock-control none -k regex:"ta::|attn_|gemm" -c 40 -o /tmp/${tag}_prof_hbm_kernels -f python tools/hbm_kernels.py > gpurun_out/${tag}_ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
python tools/ncu_summary.py /tmp/${tag}_prof_hbm_kernels.ncu-rep > gpurun_out/${tag}_prof_hbm_kernels_ncu_full.txt 2> /dev/null
fi
ls -la gpurun_out | tail -20; du -sh gpurun_out
