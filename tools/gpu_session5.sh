#!/bin/bash
# GPU session: any-stream decoder after the convergence changes (colour maps, colour-dependence taint, settled streams skip, parallel chain)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session5.log) 2>&1
export PYTHONUNBUFFERED=1
ENC_AB_FOREIGN=1 timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 0 6 2 3 5 1 || echo "ENC_AB_RC=$?"
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/fuzz_decode.py 2000 11 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -2 gpurun_out/fuzz.log
ENC_AB_FOREIGN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 130 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 8 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
ENC_AB_FOREIGN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 130 --csv --log-file gpurun_out/foreign_launches_grad2.csv bench/enc_ab qoi_b200/libqoi_gpu.so 8 4096 4096 4 3 > gpurun_out/ncu_foreign2.log 2>&1; echo "NCU_F2_RC=$?"
