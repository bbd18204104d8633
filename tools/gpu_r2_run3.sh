#!/usr/bin/env bash
# round 2, GPU visit 3: FP8 operand kind of the tensor-core matcher, front end v2 at 5 CTAs/SM (120x60 tiles)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in 1 0; do
  timeout 200 python tools/gpu_knn_mma_check.py 0 time $k > gpurun_out/knn_mma_kind$k.log 2>&1
  echo "== knn mma kind $k rc=$?"; tail -12 gpurun_out/knn_mma_kind$k.log
done
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_frontend test_gpu_orb_match test_gpu_zz_experimental test_gpu_pipeline" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error" | head -20
timeout 200 python tools/gpu_frontend_ab.py 2>&1 | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_mma_launches.csv python tools/gpu_knn_mma_prof.py 2 3 1 > gpurun_out/knn_mma_prof.log 2>&1
echo "== knn launch list rc=$?"; grep -v "^==" gpurun_out/knn_mma_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | grep knn2 | tail -8
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn2_mma_kernel -s 1 -c 1 -f -o gpurun_out/prof_knn_mma_f8_r02 python tools/gpu_knn_mma_prof.py 2 3 1 > gpurun_out/ncu_knn_full.log 2>&1
echo "== ncu knn full rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run3.json 2> gpurun_out/bench_r2_run3.err
echo "== bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run3.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['value'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:frontend_tile -s 1 -c 1 -f -o gpurun_out/prof_frontend_r02b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fe_full.log 2>&1
echo "== ncu frontend full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_run3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu list rc=$?"
