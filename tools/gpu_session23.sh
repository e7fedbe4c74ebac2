#!/bin/bash
# GPU session: final tree of round 2 -- smoke, all gpu tests, bench lines (headline, reference arm, configs[2], configs[3] share)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session23.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"; echo "SMOKE_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "BENCH_REF_RC=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
for c in 1080p_rgb_1024 8192_rgba_32; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "BENCH_${c}_RC=$?"; tail -2 gpurun_out/bench_$c.err
done
