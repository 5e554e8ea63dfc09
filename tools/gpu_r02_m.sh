#!/bin/bash
mkdir -p gpurun_out && cd /root/repo
timeout 60 python tools/attn_bench.py 1 2 256 2 64 1 2 > gpurun_out/m_smoke.txt 2>&1 || { echo "SMOKE FAILED/HUNG"; tail -3 gpurun_out/m_smoke.txt; exit 1; }
timeout 60 python tools/attn_bench.py 1 2 256 2 64 3 2 >> gpurun_out/m_smoke.txt 2>&1 || { echo "SMOKE multi FAILED/HUNG"; tail -3 gpurun_out/m_smoke.txt; exit 1; }
timeout 60 python tools/attn_bench.py 1 3 100 2 16 2 2 >> gpurun_out/m_smoke.txt 2>&1 || { echo "SMOKE ragged FAILED/HUNG"; tail -3 gpurun_out/m_smoke.txt; exit 1; }
(timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -k "attention" --timeout=60 -p no:cacheprovider -x) > gpurun_out/m_attn_tests.log 2>&1
rc=$?; echo "attention kernel tests rc=$rc"; tail -5 gpurun_out/m_attn_tests.log
[ $rc -ne 0 ] && exit 1
for shape in "8 64 256 16 64 1" "1 64 256 16 64 7" "4 64 128 8 64 1"; do timeout 60 python tools/attn_bench.py $shape 10 2>&1 | tail -1; done | tee gpurun_out/m_attn_times.txt
true
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider) > gpurun_out/m_engine_tests.log 2>&1
echo "engine tests rc=$?"; tail -4 gpurun_out/m_engine_tests.log
