#!/usr/bin/env bash
# round 2, GPU visit 16: gather with 64-bit entries / precomputed offsets / 128-register cap; control-kernel CTA size A/B
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_ba test_gpu_pipeline test_gpu_system" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -30
python tools/gpu_ba_bench.py 2>&1 | tail -4
for extra in "" "--ba-ctl-threads 256" "" "--ba-ctl-threads 256"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats $extra > gpurun_out/bench_r2_run16.json 2> gpurun_out/bench_r2_run16.err
echo "== bench $extra rc=$?"; tail -3 gpurun_out/bench_r2_run16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run16.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run16.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_bench16.log 2>&1
echo "== ncu list rc=$?"
