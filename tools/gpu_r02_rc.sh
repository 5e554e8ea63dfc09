#!/bin/bash
# activation recompute: bit-identity test, c4 with cross-attention on every modality at the full batch, c4 with the switch on
mkdir -p gpurun_out && cd /root/repo
(timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "recompute" --timeout=300 -p no:cacheprovider) > gpurun_out/rc_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/rc_tests.log; grep -n "^E " gpurun_out/rc_tests.log | head -5
timeout 900 python bench.py --all-cross --recompute --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/rc_bench_c4_allcross_b64.json 2> gpurun_out/rc_bench_allcross.err; echo "allcross rc=$?"; tail -c 400 gpurun_out/rc_bench_allcross.err
timeout 900 python bench.py --recompute --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/rc_bench_c4.json 2> gpurun_out/rc_bench_c4.err; echo "c4 rc=$?"; tail -c 300 gpurun_out/rc_bench_c4.err
for f in rc_bench_c4_allcross_b64 rc_bench_c4; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['hbm_bytes_allocated']/1e9, d['clocks']['sm_mhz'])"; done
