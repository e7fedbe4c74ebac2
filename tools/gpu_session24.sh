#!/bin/bash
# GPU session: fparse walk 0 through the shared-memory chunk ring (256-bit loads) instead of byte loads
mkdir -p gpurun_out
exec > >(tee gpurun_out/session24.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 7 0 6 3 || echo "ENC_AB_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 || echo "ENC_AB_ONE_RC=$?"
timeout 200 bench/enc_ab qoi_b200/libqoi_gpu.so 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python tools/fuzz_decode.py 4000 321 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -1 gpurun_out/fuzz.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fparse|fchain|fzero' -c 12 --csv --log-file gpurun_out/fparse_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_fparse.log 2>&1; echo "NCU_RC=$?"; grep -c fparse gpurun_out/fparse_launches.csv; grep "fparse" gpurun_out/fparse_launches.csv | tail -2 | cut -c1-40,140-260
