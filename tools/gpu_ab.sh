#!/bin/bash
# A/B session: the two look-up variants of the encoder (QOI_GPU_WALK=0/1) on the bench workload, plus parity of the variant.
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab.log) 2>&1
set -x
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py smoke; echo "SMOKE_RC=$?"
for v in 0 1; do
  QOI_GPU_WALK=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_walk$v.json 2> gpurun_out/bench_walk$v.err; echo "BENCH_WALK${v}_RC=$?"
  for d in 7 2; do
    QOI_GPU_WALK=$v timeout 300 python bench.py --steps 5 --warmup 3 --dist $d --no-cpu --e2e-steps 1 > gpurun_out/bench_walk${v}_dist$d.json 2>> gpurun_out/bench_walk$v.err; echo "BENCH_WALK${v}_DIST${d}_RC=$?"
  done
done
QOI_GPU_WALK=1 timeout 900 python -m pytest tests -x -q -m gpu --timeout=240 --timeout-method=thread > gpurun_out/pytest_walk1.log 2>&1; echo "PYTEST_WALK1_RC=$?"
tail -3 gpurun_out/pytest_walk1.log
