#!/usr/bin/env bash
# round 2, GPU visit 8: CUDA-graph pipeline, checksums, C3 bench, launch list, full ncu captures of every kernel of the step
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SKIP_BENCH=1 SKIP_NCU=1 SKIP_AB=1 FILES="test_gpu_pipeline test_gpu_ba test_gpu_orb_match" bash tools/gpu_check.sh 2>&1 | grep -E "passed|failed|rc=|Error|error" | head -20
timeout 200 python bench.py --print-checksums 2>&1 | tail -2
for cfg in c2 c3; do
  timeout 400 python bench.py --config $cfg --steps 10 --warmup 3 $( [ $cfg = c3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_r2_run8_$cfg.json 2> gpurun_out/bench_r2_run8_$cfg.err
  echo "== bench $cfg rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_run8_$cfg.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','cuda_graphs','output_sha','output_check','host_numa_cpus')}, d['roofline']['frac'], d['roofline']['launch_ms'], d['e2e']['value'], d.get('cpu_baseline'))
PY
  tail -2 gpurun_out/bench_r2_run8_$cfg.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graphs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-graphs:', d['value'], d['ms_per_step'], d['cuda_graphs'])"
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_run8_ref.json 2>> gpurun_out/bench_r2_run8_c2.err; tail -c 700 gpurun_out/bench_r2_run8_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/launches_r2_run8.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench.log 2>&1
echo "== ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn2_mma|orb_describe|scharr|orb_blur|harris_kernel|ba_chol|ba_gather|ba_linearize|ba_pairs|ba_stats|ba_pre|ba_lm|ba_backsub|ba_setup|ba_post|pyrdown|retain_best|order_keys|keys_to_points|knn2_expand|frontend_tile' -s 250 -c 80 -f -o gpurun_out/prof_step_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs > gpurun_out/ncu_step_full.log 2>&1
echo "== ncu step full rc=$?"; ls -la gpurun_out/prof_step_r02.ncu-rep
