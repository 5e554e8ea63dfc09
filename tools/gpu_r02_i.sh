#!/bin/bash
# 2 GPUs: NCCL parity tests (DP == single GPU, launcher vs reference main.py), then the c4 bench at N=2 with both exchange types
mkdir -p gpurun_out && cd /root/repo
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/i_gpus.txt
(timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -q --timeout=600 -p no:cacheprovider) > gpurun_out/i_nccl_tests.log 2>&1
echo "nccl tests rc=$?"; tail -6 gpurun_out/i_nccl_tests.log
for gd in bf16 fp32; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --grad-dtype $gd > gpurun_out/i_bench_c4_n2_$gd.json 2> gpurun_out/i_bench_c4_n2_$gd.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/i_bench_c4_n2_$gd.json').read().strip().splitlines()[-1])
    print("$gd", d['ms_per_step'], d['value'], d['exchange'].get('exposed_ms_per_step'), d['clocks'])
except Exception as e:
    print("$gd failed", e); print(open('gpurun_out/i_bench_c4_n2_$gd.err').read()[-1500:])
PY
done
