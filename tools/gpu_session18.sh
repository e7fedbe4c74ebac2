#!/bin/bash
# GPU session: final tree -- all gpu tests, decode to 3 channels (8-byte stores) vs to 4, headline bench line
mkdir -p gpurun_out
exec > >(tee gpurun_out/session18.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 200 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 6 || echo "ENC_AB_RC=$?"
ENC_AB_DEC_CH=3 timeout 200 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 6 || echo "ENC_AB_DEC3_RC=$?"
ENC_AB_DEC_CH=3 ENC_AB_FOREIGN=1 timeout 200 bench/enc_ab qoi_b200/libqoi_gpu.so 8 1000 701 3 4 0 || echo "ENC_AB_DEC3_ODD_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
