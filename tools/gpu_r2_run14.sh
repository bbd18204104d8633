#!/usr/bin/env bash
# round 2, GPU visit 14: RANSAC chunk 128 + cached sampler table (loop-closure detect time), megakernel with gpu-scope fences, refactored BA path
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_ba test_gpu_init test_gpu_loopclosure test_gpu_pipeline test_gpu_system" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -30
python tools/gpu_lc_bench.py 2>&1 | tail -4
python - <<'PY'
import subprocess,sys
# BA stand-alone: launch sequence vs one-kernel loop
src=open('tools/gpu_ba_bench.py').read().replace('for name, opt in (("gather", 0), ("dense_dmma", 1), ("gather", 0)):\n    L.alva_set_option(b"ba_dense_schur", opt)','for name, opt in (("launch sequence", 0), ("one-kernel loop", 1), ("launch sequence", 0), ("one-kernel loop", 1)):\n    L.alva_set_option(b"ba_mega", opt)')
open('/tmp/ba_ab.py','w').write(src.replace("os.path.dirname(os.path.dirname(os.path.abspath(__file__)))","'/root/repo'"))
print(subprocess.run([sys.executable,'/tmp/ba_ab.py'],capture_output=True,text=True).stdout[-900:])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_run14.json 2> gpurun_out/bench_r2_run14.err
echo "== bench rc=$?"; tail -3 gpurun_out/bench_r2_run14.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run14.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['stats'].get('ba_iterations_mean'), d['stats'].get('ba_final_over_initial_cost'))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run14.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_bench14.log 2>&1
echo "== ncu list rc=$?"
du -sh gpurun_out
