#!/bin/bash
# attention backward v3 (direct read-out stores, pipelined prologue): smoke with a short timeout first, tests, timings, timeline
mkdir -p gpurun_out && cd /root/repo
timeout 90 python tools/attn_bench.py 1 8 256 16 64 2 2 > gpurun_out/t_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -5 gpurun_out/t_smoke.txt; exit 1; }
cat gpurun_out/t_smoke.txt | tail -1
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" --timeout=120 -p no:cacheprovider) > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_tests.log
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7" "8 64 128 16 64 1" "4 256 64 8 32 1"; do timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done | tee gpurun_out/t_attn_times.txt
for dbg in 1 2; do for shape in "8 64 256 16 64 1" "1 64 256 16 64 7"; do echo -n "debug=$dbg "; TA_ATTN_DEBUG=$dbg timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done; done | tee gpurun_out/t_attn_ablation.txt
TRADEAID_B200_LIB=/root/repo/trade-aid-multimodal-transformer_b200/build/trace/libtradeaid_b200_trace.so timeout 60 python tools/attn_bwd_trace.py 1 64 256 16 64 7 200 > gpurun_out/t_trace_cross.txt 2>&1
TRADEAID_B200_LIB=/root/repo/trade-aid-multimodal-transformer_b200/build/trace/libtradeaid_b200_trace.so timeout 60 python tools/attn_bwd_trace.py 8 64 256 16 64 1 100 > gpurun_out/t_trace_self.txt 2>&1
