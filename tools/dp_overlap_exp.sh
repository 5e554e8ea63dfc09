#!/bin/bash
# 2-GPU experiment: whole-buffer all-reduce vs bucketed overlap at several SM margins. usage: tools/dp_overlap_exp.sh <tag> [N]
tag=${1:-rXX}; N=${2:-2}
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_n${N}_${name}.json 2> gpurun_out/${tag}_n${N}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_n${N}_${name}.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2),'ms', round(d['value']), 'tok/s', d.get('exchange'))
except Exception as e: print('no line', e)
PY
)"
}
run plain --no-overlap
run ov_m16 --sm-margin 16
run ov_m0 --sm-margin 0
[ "$N" = 2 ] && run ov_m32 --sm-margin 32
run ov_m16_b12 --sm-margin 16 --buckets 12
tail -5 gpurun_out/${tag}_n${N}_ov_m16.err
