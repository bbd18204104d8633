#!/usr/bin/env bash
# round 2, GPU visit 7: System backend (pinned staging, pyramid graph, fused pose chain, batch entry point)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_system test_gpu_pose test_gpu_init" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|worst" | head -40
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run7.json 2> gpurun_out/bench_r2_run7.err
echo "== bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run7.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
s=d['stats']['tracking_stages_us']['system_api']
print('480p', s.get('ms_per_tracked_frame_median'), s.get('ms_per_keyframe_median'), s.get('frames_per_sec_whole_sequence'))
for k in ('at_1280x720','at_1920x1080','concurrent_streams'): print(k, s.get(k))
PY
tail -3 gpurun_out/bench_r2_run7.err
