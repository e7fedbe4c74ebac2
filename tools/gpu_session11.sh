#!/bin/bash
# GPU session: pipelined decoder reader (3) vs op-table reader 1; any-stream decoder after the continuation fix on the configs that fell back
mkdir -p gpurun_out
exec > >(tee gpurun_out/session11.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_r1.so; do
  timeout 300 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
timeout 300 bench/enc_ab qoi_b200/libqoi_gpu.so 32 8192 8192 4 4 || echo "ENC_AB_8K_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
unset ENC_AB_FSTATS
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/fuzz_decode.py 3000 21 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -2 gpurun_out/fuzz.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec_r3 -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU_RC=$?"
