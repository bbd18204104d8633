#!/usr/bin/env bash
# round 2, GPU visit 1: tensor-core matcher bring-up (3 layout modes), antipodal front-end A/B, baseline bench line
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for m in 0 1 2; do
  timeout 200 python tools/gpu_knn_mma_check.py $m time > gpurun_out/knn_mma_$m.log 2>&1
  echo "== knn mma mode $m rc=$?"; tail -25 gpurun_out/knn_mma_$m.log
done
SKIP_BENCH=1 SKIP_NCU=1 FILES="test_gpu_zz_experimental" bash tools/gpu_check.sh 2>&1 | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run1.json 2> gpurun_out/bench_r2_run1.err
echo "== bench rc=$?"; tail -c 1500 gpurun_out/bench_r2_run1.json; tail -3 gpurun_out/bench_r2_run1.err
