#!/bin/bash
# GPU session: any-stream decoder with run-chasing walks (only the heads of runs of listed segments have a lane; it walks the run)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session21.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 0 6 2 3 5 1 || echo "ENC_AB_RC=$?"
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 8192 8192 4 4 || echo "ENC_AB_8K_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
unset ENC_AB_FSTATS
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/fuzz_decode.py 8000 91 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -2 gpurun_out/fuzz.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 200 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
