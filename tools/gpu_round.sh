#!/bin/bash
# One GPU-box session: smoke -> gpu tests -> bench -> ncu launch list -> ncu full capture.
# Everything is bounded by `timeout`; all logs go to gpurun_out/.  Usage: tools/gpu_round.sh [quick|full|prof]
MODE=${1:-full}
mkdir -p gpurun_out
exec > >(tee gpurun_out/round.log) 2>&1
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nproc; free -g | head -2
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py smoke; echo "SMOKE_RC=$?"
if [ "$MODE" != "prof" ]; then
  # exactly what the driver runs (one process); per-test limit with stack dump so that a hang is named, not silent
  timeout 1200 python -m pytest tests -x -q -m gpu --timeout=240 --timeout-method=thread --durations=5 > gpurun_out/pytest.log 2>&1; echo "PYTEST_RC=$?"
  tail -25 gpurun_out/pytest.log
fi
if [ "$MODE" = "quick" ]; then exit 0; fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "BENCH_RC=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for d in 0 7 2 6; do  # worst case, best case, run-heavy, index-heavy
  timeout 600 python bench.py --steps 5 --warmup 3 --dist $d --no-cpu --e2e-steps 1 > gpurun_out/bench_dist$d.json 2>> gpurun_out/bench.err; echo "BENCH_DIST${d}_RC=$?"
  cat gpurun_out/bench_dist$d.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "REF_RC=$?"
cat gpurun_out/bench_ref.json
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 8 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "NCU_LIST_RC=$?"
grep -c qoi gpurun_out/launches.csv
# full capture of both kernels at the bench batch (same command as the bench line, fewer steps)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:qoi_encode -s 3 -c 2 -o gpurun_out/prof_enc -f \
    python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "NCU_ENC_RC=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:qoi_decode -s 3 -c 2 -o gpurun_out/prof_dec -f \
    python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu --e2e-steps 1 >> gpurun_out/ncu_full.log 2>&1; echo "NCU_DEC_RC=$?"
# foreign-stream (any-stream) decoder: per-kernel launch list on the headline content and batch throughput per distribution
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^f' -c 80 --csv --log-file gpurun_out/foreign_launches_d4.csv python tools/foreign_prof.py 4 8 1 > gpurun_out/foreign_prof_d4.log 2>&1; echo "FPROF_RC=$?"
for d in 4 0 6 2 3 7; do timeout 200 python tools/foreign_prof.py $d 32 3 2>&1 | tail -1; done | tee gpurun_out/foreign_batch32.txt
timeout 200 python tools/foreign_prof.py 4 1 3 2>&1 | tail -1 | tee -a gpurun_out/foreign_batch32.txt
# A/B: chunks of the pipelined host path (e2e only)
for k in 8 16 32; do
  QOI_GPU_HOST_CHUNKS=$k timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 4 2>> gpurun_out/bench.err | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('host chunks $k: e2e', j['e2e'])"
done | tee gpurun_out/host_chunks.txt
ls -la gpurun_out
