#!/usr/bin/env bash
# round 2, GPU visit 12: lagged BA schedule (two chains in flight), front-end phase E in registers
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_frontend test_gpu_zz_experimental test_gpu_loopclosure test_gpu_pipeline" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -30
python tools/gpu_frontend_ab.py 2>&1 | tail -9 | head -3
for extra in "" "--no-ba-lag"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats $extra > gpurun_out/bench_r2_run12.json 2> gpurun_out/bench_r2_run12.err
echo "== bench $extra rc=$?"; tail -3 gpurun_out/bench_r2_run12.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run12.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['stats'].get('ba_iterations_mean'), d['stats'].get('ba_final_over_initial_cost'))
PY
done
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_run12_c3.json 2> gpurun_out/bench_r2_run12_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run12_c3.json').read().strip().splitlines()[-1])
print('c3',{k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
PY
du -sh gpurun_out
