#!/bin/bash
# GPU session: 256-bit loads/stores in the decoders (L1 data pipe bound): op table 16 B (default) vs 4 B (l32) vs pipelined reader (r3)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session12.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_l32.so qoi_b200/libqoi_gpu_r3.so; do
  timeout 300 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 || echo "ENC_AB_ONE_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec_256 -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec_256_l32 -f bench/enc_ab qoi_b200/libqoi_gpu_l32.so 32 4096 4096 4 4 > gpurun_out/ncu_dec2.log 2>&1; echo "NCU2_RC=$?"
