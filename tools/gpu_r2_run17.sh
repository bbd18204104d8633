#!/usr/bin/env bash
# round 2, GPU visit 17: why does one round of the five-point RANSAC kernel take 2.8 ms?  (full ncu capture with source counters)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'essential_kernel' -s 1 -c 1 -f -o gpurun_out/prof_essential python tools/gpu_lc_bench.py > gpurun_out/ncu_ess.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_essential.ncu-rep
for extra in "" "--frontend-ctas 4" "" "--frontend-ctas 4"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats $extra > gpurun_out/bench_r2_run17.json 2> gpurun_out/bench_r2_run17.err
echo "== bench $extra rc=$?"; tail -3 gpurun_out/bench_r2_run17.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run17.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check')}, d['roofline']['frac'], d['roofline'].get('launch_ms'), d['e2e']['value'])
PY
done
