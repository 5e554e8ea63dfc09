#!/bin/bash
# Round artefacts in one gpurun call: GPU tests, c4 bench line (+cpu_baseline), reference arm, ncu launch list.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_gpu_tests.log
TRADEAID_BENCH_DETAIL=gpurun_out/${tag}_gemm_detail_c4.txt timeout 900 python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_c4.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c4_reference.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${tag}_bench_c4_reference.json | cut -c1-600
timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2700 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu rc=$?"
