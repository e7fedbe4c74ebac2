#!/bin/bash
# GPU session: ncu --set full of the op-table decoder (why did 35 % fewer instructions buy nothing?), fparse with 16-byte pieces
mkdir -p gpurun_out
exec > >(tee gpurun_out/session10.log) 2>&1
export PYTHONUNBUFFERED=1
ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1 timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 6 || echo "ENC_AB_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec_lut -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU_RC=$?"
ENC_AB_FOREIGN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 200 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "foreign or sidecar or kat" --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
