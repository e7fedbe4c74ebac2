#!/bin/bash
# 8-GPU session: BASELINE configs[3] proper -- 256 x 8192x8192 RGBA over 8 GPUs through the multi-device layer
N=${1:-8}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi8.log) 2>&1
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python tools/bench_multi.py --gpus $N --per-gpu 32 --root-per-gpu 4 --reps 3 > gpurun_out/bench_multi_8192_n$N.jsonl 2> gpurun_out/bench_multi_n$N.err; echo "BENCH_MULTI_RC=$?"
cat gpurun_out/bench_multi_8192_n$N.jsonl; tail -3 gpurun_out/bench_multi_n$N.err
