#!/bin/bash
# wide bf16 epilogue stores (+ wide saved-activation loads): correctness, stage-2 timings, c4 step with the switch off / on
mkdir -p gpurun_out && cd /root/repo
(timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "wide" --timeout=150 -p no:cacheprovider) > gpurun_out/wide_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -15 gpurun_out/wide_smoke.txt | cut -c1-300; exit 1; }
tail -1 gpurun_out/wide_smoke.txt
for m in 1 2; do (TA_GEMM_WIDE=$m timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_parity_gpu.py -q -x --timeout=300 -p no:cacheprovider) > gpurun_out/wide_tests_$m.log 2>&1; echo "tests mode $m rc=$?"; tail -2 gpurun_out/wide_tests_$m.log; done
for m in 0 1; do TA_GEMM_WIDE=$m timeout 60 python tools/stage2_time.py; done | tee gpurun_out/wide_stage2.txt
for m in 0 1 0 1; do TA_GEMM_WIDE=$m TRADEAID_BENCH_DETAIL=gpurun_out/wide_gemm_detail_$m.txt timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/wide_bench_$m.json 2> gpurun_out/wide_bench_$m.err; python -c "
import json
d=json.loads(open('gpurun_out/wide_bench_$m.json').read().strip().splitlines()[-1]); print('mode $m', d['ms_per_step'], d['clocks']['sm_mhz'], d['breakdown_eager_step']['gemm']['ms'])"; done
