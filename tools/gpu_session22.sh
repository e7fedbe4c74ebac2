#!/bin/bash
# GPU session: any-stream decoder with runs of listed segments cut every 4 (QOI_FRUN)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session22.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 6 3 0 || echo "ENC_AB_RC=$?"
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
unset ENC_AB_FSTATS
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python tools/fuzz_decode.py 3000 123 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -1 gpurun_out/fuzz.log
