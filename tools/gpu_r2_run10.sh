#!/usr/bin/env bash
# round 2, GPU visit 10: BA chain rework (tiled LDL', wider gather / control kernels, forked structure kernels), loop-closure
# tests with a really unrelated scene, and the crash seen under ncu at visit 9 (faulthandler stack, with and without ncu)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_ba test_gpu_loopclosure test_gpu_pipeline test_gpu_system" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run10.json 2> gpurun_out/bench_r2_run10.err
echo "== bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run10.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, d['roofline']['frac'], d['e2e'])
print(d.get('stats',{}).get('stage_us'))
PY
MALLOC_CHECK_=3 timeout 300 python -X faulthandler bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
echo "== no-graphs bench rc=$?"; tail -30 gpurun_out/bench_nograph.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run10.csv python -X faulthandler bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench10.log 2>&1
echo "== ncu list rc=$?"; grep -v PROF gpurun_out/ncu_bench10.log | tail -40
du -sh gpurun_out
