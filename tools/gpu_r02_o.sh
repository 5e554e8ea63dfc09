#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
(timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -q --timeout=600 -p no:cacheprovider) > gpurun_out/o_nccl_tests.log 2>&1
echo "nccl tests rc=$?"; tail -6 gpurun_out/o_nccl_tests.log; grep -n "^E " gpurun_out/o_nccl_tests.log | cut -c1-600 | head
tail -3 gpurun_out/parity_r02.jsonl | cut -c1-700
