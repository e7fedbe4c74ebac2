#!/bin/bash
# GPU session: ncu --set full of encoder (photo), indexed decoder (photo), and the launch list of the any-stream decoder
mkdir -p gpurun_out
exec > >(tee gpurun_out/session7.log) 2>&1
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_enc.log 2>&1; echo "NCU_RC=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_decode -s 2 -c 1 -o gpurun_out/ncu_dec -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_dec.log 2>&1; echo "NCU3_RC=$?"
ENC_AB_FOREIGN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 200 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
ls -la gpurun_out | head -30
