#!/bin/bash
# GPU session: encoder A/B + any-stream decoder timings, full GPU test suite, fuzz, bench lines
mkdir -p gpurun_out
exec > >(tee gpurun_out/session4.log) 2>&1
export PYTHONUNBUFFERED=1
for lib in qoi_b200/libqoi_gpu.so qoi_b200/libqoi_gpu_mb8.so; do
  ENC_AB_FOREIGN=1 timeout 200 bench/enc_ab $lib 32 4096 4096 4 4 7 0 6 2 3 5 || echo "ENC_AB_RC=$?"
done
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 4096 4096 4 4 7 6 || echo "ENC_AB_ONE_RC=$?"
ENC_AB_FOREIGN=1 timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 64 1920 1080 3 4 7 || echo "ENC_AB_RGB_RC=$?"
timeout 120 bench/enc_ab qoi_b200/libqoi_gpu.so 1 16384 16384 4 4 7 || echo "ENC_AB_BIG_RC=$?"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=300 --timeout-method=thread --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "PYTEST_RC=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/fuzz_decode.py 3000 7 65536 > gpurun_out/fuzz.log 2>&1; echo "FUZZ_RC=$?"; tail -2 gpurun_out/fuzz.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qoi_encode -s 2 -c 1 -o gpurun_out/ncu_enc -f bench/enc_ab qoi_b200/libqoi_gpu.so 32 4096 4096 4 4 > gpurun_out/ncu_enc.log 2>&1; echo "NCU_RC=$?"
ENC_AB_FOREIGN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f|qoi_decode' -c 120 --csv --log-file gpurun_out/foreign_launches.csv bench/enc_ab qoi_b200/libqoi_gpu.so 8 4096 4096 4 4 > gpurun_out/ncu_foreign.log 2>&1; echo "NCU_F_RC=$?"
ls -la gpurun_out | head -30
