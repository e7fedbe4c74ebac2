#!/bin/bash
# A/B of the decoders' stream-reader refill (QOI_GPU_REFILL=uniform|lane): parity in both modes, then the bench lines
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab.log) 2>&1
set -x
export PYTHONUNBUFFERED=1
for m in uniform lane; do
  QOI_GPU_REFILL=$m timeout 600 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_$m.log 2>&1; echo "PYTEST_${m}_RC=$?"
  tail -2 gpurun_out/pytest_$m.log
done
for m in uniform lane; do
  for d in 4 0 6; do
    QOI_GPU_REFILL=$m timeout 300 python bench.py --steps 10 --warmup 3 --dist $d --no-cpu --e2e-steps 1 > gpurun_out/ab_${m}_d$d.json 2>> gpurun_out/ab.err; echo "BENCH_${m}_${d}_RC=$?"
  done
done
python tools/bench_summary.py gpurun_out/ab_*.json
