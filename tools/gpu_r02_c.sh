#!/bin/bash
# ncu --set full of the two-stream attention forward at the c4 self shape and at the long-context shape
mkdir -p gpurun_out && cd /root/repo
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -s 2 -c 1 -f -o gpurun_out/c_fwd2_self python tools/attn_fwd_only.py 8 64 256 16 64 1 3 > gpurun_out/c_ncu_self.log 2>&1; tail -3 gpurun_out/c_ncu_self.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -s 2 -c 1 -f -o gpurun_out/c_fwd2_long python tools/attn_fwd_only.py 4 32 1024 12 64 1 3 > gpurun_out/c_ncu_long.log 2>&1; tail -3 gpurun_out/c_ncu_long.log
ls -la gpurun_out/*.ncu-rep | tail -3
