#!/bin/bash
# GPU session: C-API latency with stage trace, compute-sanitizer (memcheck / racecheck) over a parity subset + fuzz, bench lines of the other configs
mkdir -p gpurun_out
exec > >(tee gpurun_out/session8.log) 2>&1
export PYTHONUNBUFFERED=1
L=qoi_b200/libqoi_gpu.so
cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null; nproc
echo "== capi latency"
QOI_GPU_TRACE=1 timeout 200 python tools/capi_lat.py 4096 4096 4 4 6 2>&1 | tail -16
timeout 200 python tools/capi_lat.py 1920 1080 3 4 8 2>&1 | tail -1
QOI_GPU_HOST_THREADS=32 timeout 200 python tools/capi_lat.py 4096 4096 4 4 6 2>&1 | tail -1
echo "== memcheck enc_ab"
ENC_AB_FOREIGN=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 bench/enc_ab $L 6 1000 700 4 4 7 0 6 2 3 > gpurun_out/memcheck_enc_ab_rgba.log 2>&1; echo "MEMCHECK_RGBA_RC=$?"; tail -4 gpurun_out/memcheck_enc_ab_rgba.log
ENC_AB_FOREIGN=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 bench/enc_ab $L 6 1000 700 3 4 6 > gpurun_out/memcheck_enc_ab_rgb.log 2>&1; echo "MEMCHECK_RGB_RC=$?"; tail -4 gpurun_out/memcheck_enc_ab_rgb.log
echo "== memcheck parity subset"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kat or reject or tiny or foreign_streams or misaligned or header_mismatch or long_runs" > gpurun_out/memcheck_pytest.log 2>&1; echo "MEMCHECK_PYTEST_RC=$?"; tail -6 gpurun_out/memcheck_pytest.log
echo "== memcheck fuzz"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/fuzz_decode.py 300 3 65536 > gpurun_out/memcheck_fuzz.log 2>&1; echo "MEMCHECK_FUZZ_RC=$?"; tail -4 gpurun_out/memcheck_fuzz.log
echo "== racecheck enc_ab"
ENC_AB_FOREIGN=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 bench/enc_ab $L 4 640 480 4 4 6 > gpurun_out/racecheck_enc_ab.log 2>&1; echo "RACECHECK_RC=$?"; tail -8 gpurun_out/racecheck_enc_ab.log
echo "== fuzz parity 10k"
timeout 900 python tools/fuzz_decode.py 10000 7 65536 > gpurun_out/fuzz10k.log 2>&1; echo "FUZZ_RC=$?"; tail -2 gpurun_out/fuzz10k.log
echo "== other configs"
for c in 1080p_rgb_1024 8192_rgba_32; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "BENCH_${c}_RC=$?"; cut -c1-1500 gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
done
timeout 1200 python bench.py --config sweep --steps 3 --warmup 3 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "BENCH_SWEEP_RC=$?"; wc -l gpurun_out/bench_sweep.json; tail -2 gpurun_out/bench_sweep.err
