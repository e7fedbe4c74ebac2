#!/bin/bash
# the round-end sequence as the driver runs it: smoke, gpu tests, reference arm, bench
mkdir -p gpurun_out
exec > >(tee gpurun_out/final.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"; echo "SMOKE_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "BENCH_REF_RC=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cut -c1-260 gpurun_out/bench.json
