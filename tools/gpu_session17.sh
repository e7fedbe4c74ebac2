#!/bin/bash
# GPU session: driver-style validation of the current tree -- smoke, all gpu tests, bench lines of every config, reference arm, C-API latency, launch list of a bench step
mkdir -p gpurun_out
exec > >(tee gpurun_out/session17.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"; echo "SMOKE_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -10 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "BENCH_REF_RC=$?"
cut -c1-600 gpurun_out/bench_ref.json
for c in 1080p_rgb_1024 8192_rgba_32; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "BENCH_${c}_RC=$?"; tail -2 gpurun_out/bench_$c.err
done
timeout 1200 python bench.py --config sweep --steps 3 --warmup 3 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "BENCH_SWEEP_RC=$?"; wc -l gpurun_out/bench_sweep.json
for d in 0 6 7; do
  timeout 600 python bench.py --steps 5 --warmup 3 --dist $d --no-cpu --no-pcie --e2e-steps 1 > gpurun_out/bench_dist$d.json 2>> gpurun_out/bench.err; echo "BENCH_DIST${d}_RC=$?"
done
QOI_GPU_TRACE=1 timeout 200 python tools/capi_lat.py 4096 4096 4 4 6 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-pcie --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "NCU_LAUNCHES_RC=$?"
