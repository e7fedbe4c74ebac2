#!/bin/bash
# GPU session: encoder slot match -- six ballots (default) vs match.any on 1, 2, 3 rows of every four; new any-stream regression test; full gpu tests
mkdir -p gpurun_out
exec > >(tee gpurun_out/session16.log) 2>&1
export PYTHONUNBUFFERED=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_m1.so qoi_b200/libqoi_gpu_m2.so qoi_b200/libqoi_gpu_m3.so; do
  timeout 300 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu_m2.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc_m2 -f bench/enc_ab qoi_b200/libqoi_gpu_m2.so 32 4096 4096 4 4 > gpurun_out/ncu_enc.log 2>&1; echo "NCU2_RC=$?"
