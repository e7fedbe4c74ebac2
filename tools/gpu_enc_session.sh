#!/bin/bash
# one GPU session for the encoder work: A/B of library builds (bench/enc_ab), GPU parity subset, ncu capture
mkdir -p gpurun_out
exec > >(tee gpurun_out/enc_session.log) 2>&1
export PYTHONUNBUFFERED=1
for lib in "$@"; do
  timeout 120 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 16384 16384 4 4 7 || echo "ENC_AB_BIG_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 || echo "ENC_AB_ONE_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_parity.log 2>&1; echo "PYTEST_RC=$?"
tail -5 gpurun_out/pytest_parity.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_enc.log 2>&1; echo "NCU_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc_const -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 7 > gpurun_out/ncu_enc_const.log 2>&1; echo "NCU2_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU3_RC=$?"
ls -la gpurun_out | head -30
