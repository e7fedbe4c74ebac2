#!/bin/bash
# GPU session: bench.py after the cpu_baseline change, both arms, as the driver runs them
mkdir -p gpurun_out
exec > >(tee gpurun_out/session19.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "BENCH_REF_RC=$?"
cut -c1-500 gpurun_out/bench_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "BENCH_DEFAULT_RC=$?"
