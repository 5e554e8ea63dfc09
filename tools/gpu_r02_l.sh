#!/bin/bash
# 8 GPUs: c4 with the bf16 bucket exchange (and fp32 for comparison), c3 (BASELINE.json configs[2] is quoted at 8 GPUs)
mkdir -p gpurun_out && cd /root/repo
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/l_gpus.txt
run() { # name, extra args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 "$@" > gpurun_out/l_$name.json 2> gpurun_out/l_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/l_$name.json').read().strip().splitlines()[-1])
    print("$name", round(d['ms_per_step'],2), round(d['value']), (d.get('exchange') or {}).get('exposed_ms_per_step'), d['clocks'])
except Exception as e:
    print("$name failed", e); print(open('gpurun_out/l_$name.err').read()[-800:])
PY
}
run bench_c4_n8_bf16 --steps 10 --warmup 3 --grad-dtype bf16
run bench_c4_n8_fp32 --steps 10 --warmup 3 --grad-dtype fp32
run bench_c3_n8_bf16 --config c3 --steps 20 --warmup 5 --grad-dtype bf16
