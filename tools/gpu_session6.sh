#!/bin/bash
# GPU session: state of play after re-entry -- kernel times for all content classes, full gpu test suite, headline bench line
mkdir -p gpurun_out
exec > >(tee gpurun_out/session6.log) 2>&1
export PYTHONUNBUFFERED=1
ENC_AB_FOREIGN=1 timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 0 6 2 3 5 1 || echo "ENC_AB_RC=$?"
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -16 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
