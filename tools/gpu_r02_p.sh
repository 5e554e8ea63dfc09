#!/bin/bash
# probe: attention backward ablations + CTA timeline + line-level ncu; stage-2 GEMM ncu
mkdir -p gpurun_out && cd /root/repo
timeout 120 python tools/attn_bench.py 1 8 256 16 64 2 2 > gpurun_out/p_smoke.txt 2>&1 || { echo "smoke failed"; tail -5 gpurun_out/p_smoke.txt; exit 1; }
for dbg in 0 1 2 4 7; do for shape in "8 64 256 16 64 1" "1 64 256 16 64 7"; do echo -n "debug=$dbg "; TA_ATTN_DEBUG=$dbg timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done; done | tee gpurun_out/p_attn_ablation.txt
TRADEAID_B200_LIB=/root/repo/trade-aid-multimodal-transformer_b200/build/trace/libtradeaid_b200_trace.so timeout 60 python tools/attn_bwd_trace.py 1 64 256 16 64 7 200 > gpurun_out/p_trace_cross.txt 2>&1
TRADEAID_B200_LIB=/root/repo/trade-aid-multimodal-transformer_b200/build/trace/libtradeaid_b200_trace.so timeout 60 python tools/attn_bwd_trace.py 8 64 256 16 64 1 100 > gpurun_out/p_trace_self.txt 2>&1
tail -3 gpurun_out/p_trace_cross.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/p_bwd_cross python tools/attn_bwd_only.py 1 64 256 16 64 7 3 > gpurun_out/p_ncu_bwd_cross.log 2>&1; tail -1 gpurun_out/p_ncu_bwd_cross.log
timeout 60 python tools/stage2_time.py | tee gpurun_out/p_stage2.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gemm -s 6 -c 1 -f -o gpurun_out/p_stage2_fwd python tools/stage2_time.py > gpurun_out/p_ncu_s2f.log 2>&1; tail -1 gpurun_out/p_ncu_s2f.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gemm -s 20 -c 1 -f -o gpurun_out/p_stage2_bwd python tools/stage2_time.py > gpurun_out/p_ncu_s2b.log 2>&1; tail -1 gpurun_out/p_ncu_s2b.log
ls -la gpurun_out | grep " p_"
