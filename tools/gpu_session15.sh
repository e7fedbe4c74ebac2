#!/bin/bash
# GPU session: decoder op table 16 B (default) vs 8 B (l64); encoder slot match by six ballots (default) vs one match.any per row (match)
mkdir -p gpurun_out
exec > >(tee gpurun_out/session15.log) 2>&1
export PYTHONUNBUFFERED=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_l64.so qoi_b200/libqoi_gpu_match.so; do
  timeout 300 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu_l64.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu_match.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec_l64 -f bench/enc_ab qoi_b200/libqoi_gpu_l64.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc_match -f bench/enc_ab qoi_b200/libqoi_gpu_match.so 32 4096 4096 4 4 > gpurun_out/ncu_enc.log 2>&1; echo "NCU2_RC=$?"
