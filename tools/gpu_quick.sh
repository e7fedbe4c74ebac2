#!/bin/bash
# short session: gpu tests (driver style) + the headline bench line
mkdir -p gpurun_out
exec > >(tee gpurun_out/quick.log) 2>&1
set -x
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py smoke; echo "SMOKE_RC=$?"
timeout 1200 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread --durations=6 > gpurun_out/pytest.log 2>&1; echo "PYTEST_RC=$?"
tail -12 gpurun_out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for d in 0 6; do
  timeout 600 python bench.py --steps 5 --warmup 3 --dist $d --no-cpu --e2e-steps 1 > gpurun_out/bench_dist$d.json 2>> gpurun_out/bench.err; echo "BENCH_DIST${d}_RC=$?"
done
# per-kernel times of the foreign-stream pipeline (launch list; numbers under ncu are not bench values)
for d in 4 6; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f' -c 80 --csv --log-file gpurun_out/foreign_launches_d$d.csv python tools/foreign_prof.py $d 8 1 > gpurun_out/foreign_prof_d$d.log 2>&1; echo "FPROF${d}_RC=$?"
done
for d in 4 0 6; do
  timeout 200 python tools/foreign_prof.py $d 32 3 2>&1 | tail -1
done
