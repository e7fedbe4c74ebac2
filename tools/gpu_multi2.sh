#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the multi-device layer on hardware -- bytes at N GPUs == bytes at 1 GPU (tests/test_multi_gpu.py),
# BASELINE configs[3] (32 x 8192^2 RGBA per GPU) device-resident and incl. NCCL scatter/gather, then bench.py as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi.log) 2>&1
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu --timeout=600 --timeout-method=thread 2>&1 | tail -5; echo "PYTEST_MULTI_RC=${PIPESTATUS[0]}"
timeout 900 python tools/bench_multi.py --gpus $N --per-gpu 32 --root-per-gpu 8 --reps 4 > gpurun_out/bench_multi_8192.jsonl 2> gpurun_out/bench_multi.err; echo "BENCH_MULTI_RC=$?"
cat gpurun_out/bench_multi_8192.jsonl; tail -3 gpurun_out/bench_multi.err
timeout 600 python tools/bench_multi.py --gpus $N --w 4096 --h 4096 --per-gpu 32 --root-per-gpu 16 --reps 4 > gpurun_out/bench_multi_4096.jsonl 2>> gpurun_out/bench_multi.err; echo "BENCH_MULTI4K_RC=$?"
cat gpurun_out/bench_multi_4096.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "BENCHN_RC=$?"
cut -c1-1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
