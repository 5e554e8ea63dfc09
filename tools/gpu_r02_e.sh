#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
timeout 60 python tools/attn_fwd_only.py 1 2 256 2 64 1 2 > gpurun_out/e_smoke.txt 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -3 gpurun_out/e_smoke.txt; exit 1; }
timeout 60 python tools/attn_fwd_only.py 1 2 256 2 64 3 2 >> gpurun_out/e_smoke.txt 2>&1 || { echo "SMOKE (multi) FAILED/HUNG"; tail -3 gpurun_out/e_smoke.txt; exit 1; }
(timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -k "attention" --timeout=60 -p no:cacheprovider -x) > gpurun_out/e_attn_tests.log 2>&1
rc=$?; echo "attention kernel tests rc=$rc"; tail -5 gpurun_out/e_attn_tests.log
[ $rc -ne 0 ] && exit 1
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7" "4 64 128 8 64 1" "4 32 1024 12 64 1"; do
  timeout 60 python tools/attn_fwd_only.py $shape 10 2>&1 | tail -1 | sed "s/^/v2 /"
done | tee gpurun_out/e_attn_fwd_times.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -s 2 -c 1 -f -o gpurun_out/e_fwd2_self python tools/attn_fwd_only.py 8 64 256 16 64 1 3 > gpurun_out/e_ncu_self.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -s 2 -c 1 -f -o gpurun_out/e_fwd2_cross python tools/attn_fwd_only.py 1 64 256 16 64 7 3 > gpurun_out/e_ncu_cross.log 2>&1
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider) > gpurun_out/e_engine_tests.log 2>&1
echo "engine tests rc=$?"; tail -4 gpurun_out/e_engine_tests.log
