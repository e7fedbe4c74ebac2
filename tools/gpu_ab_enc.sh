#!/bin/bash
# Round-2 opener: A/B of the encoder variants that are off by default (cuda/qoi_enc4.cuh, experiments/).
#   stage 1 (seconds)   bench/striped_smoke 32: every variant against the shipped kernel, byte-identical streams + CUDA-event times
#   stage 2 (per variant named on the command line; default: none) GPU parity tests that exercise the encoder, bench lines for
#                       photo / noise / palette / const0, one ncu --set full capture
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab_enc.sh slab-match striped-match'
# QOI_GPU_DEC=half (two decode lanes per tile, needs a slab* encoder) can be exported around the call; stage 1 times it as *+half
# Variants: striped | striped-match | striped-earlyw | slab | slab-match | slab-pipe | slab-pipe-ballot | slab4 | slab4-pipe   (QOI_GPU_ENC), earlyw (needs qoi_b200/libqoi_gpu_exp.so)
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab_enc.log) 2>&1
export PYTHONUNBUFFERED=1
EXP=$PWD/qoi_b200/libqoi_gpu_exp.so
echo "=== stage 1: striped_smoke"
timeout 120 bench/striped_smoke 32 | tee gpurun_out/striped_smoke.txt
echo "SMOKE_RC=$?"
if [ -f "$EXP" ]; then  # the patched shipped encoder (experiments/series), same tool with the experimental library in front
  LD_PRELOAD=$EXP timeout 120 bench/striped_smoke 32 earlyw mid+half earlyw-mid+half | tee gpurun_out/striped_smoke_exp.txt
  echo "SMOKE_EXP_RC=$?"
fi
echo "=== stage 1b: parity of every off-by-default variant (mixed batches, RGB/RGBA, misaligned, partial tiles), one process per case"
QOI_TEST_VARIANTS=1 timeout 1200 python -m pytest tests/test_zz_variants.py -q -m gpu 2>&1 | tail -5
run_variant() {  # name, env assignments...
  local name=$1; shift
  echo "=== $name"
  env "$@" timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=300 --timeout-method=thread \
      -k "not foreign and not photograph" > gpurun_out/ab_enc_pytest_$name.log 2>&1
  echo "PYTEST_${name}_RC=$?"; tail -2 gpurun_out/ab_enc_pytest_$name.log
  for d in 4 0 6 7; do
    env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --dist $d --no-cpu --e2e-steps 1 > gpurun_out/ab_enc_${name}_d$d.json 2>> gpurun_out/ab_enc.err
    echo "BENCH_${name}_d${d}_RC=$?"
  done
  env "$@" timeout 600 ncu --set full --import-source on --clock-control none -k regex:encode -c 1 \
    -o gpurun_out/ncu_$name python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --batch 8 > gpurun_out/ncu_$name.log 2>&1
  echo "NCU_${name}_RC=$?"
}
[ $# -gt 0 ] && run_variant default QOI_GPU_NOP=1
for v in "$@"; do
  if [ "$v" = earlyw ]; then
    [ -f "$EXP" ] && run_variant earlyw QOI_GPU_LIB=$EXP QOI_GPU_ENC=earlyw
  else
    run_variant "${v//-/}" QOI_GPU_ENC=$v
  fi
done
ls gpurun_out/ab_enc_*_d*.json >/dev/null 2>&1 && python tools/bench_summary.py gpurun_out/ab_enc_*_d*.json
