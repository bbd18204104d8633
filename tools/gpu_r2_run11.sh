#!/usr/bin/env bash
# round 2, GPU visit 11: front-end v2 edits (3-D grid, per-warp emission, pyramid shuffles, L2 prefetch A/B), BA gather split +
# fast pivot reciprocal, loop-closure test; BA stand-alone time, bench with / without the BA overlap
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_frontend test_gpu_zz_experimental test_gpu_ba test_gpu_loopclosure test_gpu_pipeline test_gpu_system test_gpu_orb_match" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -30
python tools/gpu_frontend_ab.py 2>&1 | tail -12
python tools/gpu_ba_bench.py 2>&1 | tail -6
for extra in "" "--no-ba-overlap"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats $extra > gpurun_out/bench_r2_run11.json 2> gpurun_out/bench_r2_run11.err
echo "== bench $extra rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run11.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, d['roofline']['frac'], d['roofline'].get('launch_us'))
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run11.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_bench11.log 2>&1
echo "== ncu list rc=$?"; grep -v PROF gpurun_out/ncu_bench11.log | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'frontend_tile_kernel_v2' -s 3 -c 1 -f -o gpurun_out/prof_frontend_r02c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_fe11.log 2>&1
echo "== ncu frontend rc=$?"
du -sh gpurun_out
