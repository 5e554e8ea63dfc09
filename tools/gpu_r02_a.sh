#!/bin/bash
# round 2, GPU call A: the whole -m gpu suite (incl. c4-width parity, bf16x3, main.py on cuda, generate vs reference), then benches
mkdir -p gpurun_out && cd /root/repo
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpus.txt
(time timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err; tail -c 600 gpurun_out/a_bench_c4.err
timeout 600 python bench.py --tier compat --steps 5 --warmup 3 > gpurun_out/a_bench_c4_compat.json 2> gpurun_out/a_bench_c4_compat.err; tail -c 400 gpurun_out/a_bench_c4_compat.err
timeout 600 python bench.py --precision bf16x3 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c4_bf16x3_b16.json 2> gpurun_out/a_bench_c4_bf16x3.err; tail -c 400 gpurun_out/a_bench_c4_bf16x3.err
timeout 300 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c4_bf16_b16.json 2>/dev/null
for mode in fp32 tf32 autocast; do
  timeout 300 python bench.py --impl reference-eager --config c2 --ref-mode $mode --steps 5 --warmup 2 > gpurun_out/a_eager_c2_$mode.json 2> gpurun_out/a_eager_c2_$mode.err
  timeout 400 python bench.py --impl reference-eager --config c3 --ref-mode $mode --steps 3 --warmup 1 > gpurun_out/a_eager_c3_$mode.json 2> gpurun_out/a_eager_c3_$mode.err
done
timeout 900 python bench.py --impl reference-eager --config c4 --batch 4 --ref-mode autocast --steps 2 --warmup 1 > gpurun_out/a_eager_c4_b4_autocast.json 2> gpurun_out/a_eager_c4.err
timeout 300 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_c2.json 2>/dev/null
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_c3.json 2>/dev/null
ls -la gpurun_out | tail -30
