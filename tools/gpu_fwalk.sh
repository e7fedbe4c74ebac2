#!/bin/bash
# foreign-stream decoder: parity tests + per-kernel times + batch-32 throughput per distribution
mkdir -p gpurun_out
exec > >(tee gpurun_out/fwalk.log) 2>&1
set -x
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest.log
for d in 4 6 0; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f' -c 80 --csv --log-file gpurun_out/foreign_launches_d$d.csv python tools/foreign_prof.py $d 8 1 > gpurun_out/foreign_prof_d$d.log 2>&1; echo "FPROF_${d}_RC=$?"
done
for d in 4 0 6 2 3 7; do timeout 200 python tools/foreign_prof.py $d 32 3 2>&1 | tail -1; done
timeout 200 python tools/foreign_prof.py 4 1 3 2>&1 | tail -1
