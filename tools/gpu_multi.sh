#!/bin/bash
# Multi-GPU session (gpurun --gpus N): gpu tests once, then bench.py at 1 and N GPUs exactly as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi.log) 2>&1
set -x
nvidia-smi --query-gpu=index,name,memory.total --format=csv
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "PYTEST_RC=${PIPESTATUS[0]}"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "BENCH1_RC=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "BENCHN_RC=$?"
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; echo "REFN_RC=$?"
cat gpurun_out/bench_ref_n$N.json
