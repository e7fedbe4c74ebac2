#!/bin/bash
# GPU session: state rounds with 256-bit table I/O; fparse back to byte loads; launch list + ncu of the first state round
mkdir -p gpurun_out
exec > >(tee gpurun_out/session14.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
unset ENC_AB_FSTATS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 200 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fstate -s 16 -c 1 -o gpurun_out/ncu_fstate -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_fstate.log 2>&1; echo "NCU_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
