#!/usr/bin/env bash
# round 2, GPU visit 9: loop-closure tests, launch list, full ncu captures of the step's kernels (text summaries made on the box)
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_loopclosure test_gpu_pipeline test_gpu_ba test_gpu_system" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error|assert" | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_run9.json 2> gpurun_out/bench_r2_run9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_run9.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','output_check')}, d['roofline']['frac'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run9.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench.log 2>&1
echo "== ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'knn2_mma|orb_describe|scharr|orb_blur|harris_kernel|ba_chol|ba_gather|ba_linearize|ba_pairs|ba_stats|ba_pre|ba_lm|ba_backsub|ba_setup|ba_post|pyrdown|retain_best|order_keys|keys_to_points|knn2_expand|frontend_tile' -s 250 -c 70 -f -o /tmp/prof_step_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_step_full.log 2>&1
echo "== ncu step full rc=$?"
python tools/ncu_rep_to_text.py /tmp/prof_step_r02.ncu-rep > gpurun_out/r02_kernels_full.txt 2> gpurun_out/ncu_to_text.err; wc -l gpurun_out/r02_kernels_full.txt
# the dense (FP64 tensor-core) Schur variant: tensor-pipe utilisation
timeout 300 ncu --set full --clock-control none -k regex:'ba_syrk_dmma' -c 2 -f -o /tmp/prof_dmma python tools/gpu_ba_bench.py dense > gpurun_out/ncu_dmma.log 2>&1
python tools/ncu_rep_to_text.py /tmp/prof_dmma.ncu-rep > gpurun_out/r02_dmma_full.txt 2>> gpurun_out/ncu_to_text.err; tail -25 gpurun_out/r02_dmma_full.txt
du -sh gpurun_out
