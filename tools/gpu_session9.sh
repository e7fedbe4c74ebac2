#!/bin/bash
# GPU session: decoder op-table (LUT) build, reader 1 vs 2, continuation limit old vs new on the configs where the any-stream decoder fell off a cliff
mkdir -p gpurun_out
exec > >(tee gpurun_out/session9.log) 2>&1
export PYTHONUNBUFFERED=1 ENC_AB_FOREIGN=1 ENC_AB_FSTATS=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_r2.so; do
  timeout 300 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 || echo "ENC_AB_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu_oldcont.so 32 4096 4096 4 4 3 || echo "ENC_AB_RC=$?"
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_oldcont.so; do
  timeout 300 bench/enc_ab $lib 8 8192 8192 4 4 || echo "ENC_AB_8K_RC=$?"
  timeout 300 bench/enc_ab $lib 512 1920 1080 3 4 || echo "ENC_AB_1080_RC=$?"
done
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 3 || echo "ENC_AB_ONE_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu_r2.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
unset ENC_AB_FSTATS
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 bench/enc_ab qoi_b200/libqoi_gpu.so 2 640 480 4 4 6 > gpurun_out/racecheck_enc_ab.log 2>&1; echo "RACECHECK_RC=$?"; tail -5 gpurun_out/racecheck_enc_ab.log
