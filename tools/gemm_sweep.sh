#!/bin/bash
# Dev sweep: each case in its own process under a timeout so one hang cannot take the call down.
out=gpurun_out/gemm_sweep.jsonl; : > $out
run() { timeout 120 python tools/gemm_check.py "$@" >> $out 2>> gpurun_out/gemm_sweep.err; echo "rc=$? $*" >> gpurun_out/gemm_sweep.rc; }
: > gpurun_out/gemm_sweep.rc; : > gpurun_out/gemm_sweep.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> gpurun_out/gemm_sweep.rc
run 128 64 64 1 1 0 0
run 128 128 64 1 1 0 0
run 128 256 64 1 1 0 0
run 128 256 256 1 1 0 0
run 256 512 512 1 1 0 0
run 100 72 40 1 1 0 0
run 16 28 32 2 1 0 0
run 128 64 64 1 1 0 1
run 128 64 64 1 1 1 0
run 128 64 64 1 1 1 1
run 256 512 512 1 1 0 1
run 256 512 512 1 1 1 1
run 200 136 72 3 2 0 1
run 200 136 72 3 2 1 1
run 300 200 136 2 3 0 0 bias_tanh
run 300 200 136 2 1 0 0 bias_relu
run 300 200 136 1 1 0 0 res
run 300 200 136 1 1 0 1 tanh_bwd
run 300 200 136 1 1 0 1 relu_bwd
run 4096 4096 4096 1 1 0 0
run 4096 4096 4096 1 1 0 1
run 4096 4096 4096 1 1 1 1
run 16384 4096 1024 1 1 0 0 bias_relu
run 16384 1024 4096 1 1 0 0 res
run 16384 1536 1024 8 1 0 0 bias_tanh
run 4096 1024 16384 1 1 1 1
cat gpurun_out/gemm_sweep.rc
