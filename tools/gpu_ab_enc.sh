#!/bin/bash
# Round-2 opener (one gpurun call, ~10 GPU-minutes): A/B of the encoder variants that so far only ran on the CPU emulator.
#   default                       the shipped kernels (cuda/qoi_enc.cuh)
#   QOI_GPU_ENC=striped           lane-per-pixel encoder, six-ballot slot matching (cuda/qoi_enc4.cuh)
#   QOI_GPU_ENC=striped-match     same with match.any
#   QOI_GPU_ENC=slab | slab-match rolled loops, op bytes straight into per-warp slabs, copy-out by the compute warps
#   exp library + QOI_GPU_ENC=earlyw   phase W before the output offset is known (experiments/encoder_earlyw.patch)
# For every variant: the GPU parity tests that exercise the encoder, then bench lines for photo / noise / palette / const0.
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab_enc.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab_enc.log) 2>&1
export PYTHONUNBUFFERED=1
EXP=$PWD/qoi_b200/libqoi_gpu_exp.so
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
}
run_variant default QOI_GPU_NOP=1
run_variant striped QOI_GPU_ENC=striped
run_variant stripedmatch QOI_GPU_ENC=striped-match
run_variant stripedearlyw QOI_GPU_ENC=striped-earlyw
run_variant slab QOI_GPU_ENC=slab
run_variant slabmatch QOI_GPU_ENC=slab-match
if [ -f "$EXP" ]; then run_variant earlyw QOI_GPU_LIB=$EXP QOI_GPU_ENC=earlyw; fi
python tools/bench_summary.py gpurun_out/ab_enc_*_d*.json
# one ncu capture of the striped kernel for the line profile (tools/ncu_lines.py reads it here)
for v in striped slab; do
QOI_GPU_ENC=$v timeout 600 ncu --set full --import-source on --clock-control none -k regex:striped -c 1 \
  -o gpurun_out/ncu_$v python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --batch 8 > gpurun_out/ncu_$v.log 2>&1
echo "NCU_${v}_RC=$?"
done
