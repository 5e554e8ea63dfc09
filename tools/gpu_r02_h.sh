#!/bin/bash
# compute-sanitizer evidence on the small-shape launches of tools/sanitize_small.py
mkdir -p gpurun_out && cd /root/repo
timeout 120 python tools/sanitize_small.py > gpurun_out/h_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/h_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/h_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/h_sanitizer_$tool.log | tail -5
done
TA_ATTN_FWD=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/h_sanitizer_memcheck_fwd1.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/h_sanitizer_memcheck_fwd1.log
