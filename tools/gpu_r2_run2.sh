#!/usr/bin/env bash
# round 2, GPU visit 2: parity of the new front end + tensor-core matcher, A/B timings, per-kernel times, full ncu captures
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_frontend test_gpu_orb_match test_gpu_zz_experimental test_gpu_pipeline" bash tools/gpu_check.sh 2>&1 | tail -70
timeout 200 python tools/gpu_frontend_ab.py 2>&1 | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_mma_launches.csv python tools/gpu_knn_mma_prof.py 2 3 > gpurun_out/knn_mma_prof.log 2>&1
echo "== knn launch list rc=$?"; grep -v "^==" gpurun_out/knn_mma_launches.csv | awk -F'","' '{print $5, $NF}' | tail -14
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn2_mma_kernel -s 1 -c 1 -f -o gpurun_out/prof_knn_mma_r02 python tools/gpu_knn_mma_prof.py 2 3 > gpurun_out/ncu_knn_full.log 2>&1
echo "== ncu knn full rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run2.json 2> gpurun_out/bench_r2_run2.err
echo "== bench rc=$?"; tail -c 600 gpurun_out/bench_r2_run2.json | head -c 300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['value'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:frontend_tile -s 1 -c 1 -f -o gpurun_out/prof_frontend_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fe_full.log 2>&1
echo "== ncu frontend full rc=$?"; ls -la gpurun_out/*.ncu-rep
