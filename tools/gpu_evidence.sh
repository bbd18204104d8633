#!/usr/bin/env bash
# evidence visit (round 2): whole GPU suite, smoke, bench lines (c2 / c3 / reference arm), launch list, per-kernel ncu summaries
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -5 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "== bench rc=$?"; tail -3 gpurun_out/r02_bench.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-stage-stats > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
echo "== bench c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "== bench reference rc=$?"
python - <<'PY'
import json
for f in ('r02_bench','r02_bench_c3','r02_bench_reference'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_bench_final.log 2>&1
echo "== ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'knn2_mma|orb_describe|scharr|orb_blur|harris_kernel|ba_chol|ba_gather|ba_linearize|ba_pairs|ba_stats|ba_pre|ba_lm|ba_backsub|ba_setup|ba_post|pyrdown|retain_best|order_keys|keys_to_points|knn2_expand|frontend_tile' -s 250 -c 70 -f -o /tmp/prof_step_r02 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_step_full.log 2>&1
echo "== ncu step full rc=$?"
python tools/ncu_rep_to_text.py /tmp/prof_step_r02.ncu-rep > gpurun_out/r02_kernels_full.txt 2> gpurun_out/ncu_to_text.err; wc -l gpurun_out/r02_kernels_full.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'frontend_tile_kernel_v2' -s 3 -c 1 -f -o gpurun_out/prof_frontend_r02_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graphs --no-stage-stats > gpurun_out/ncu_fe_final.log 2>&1
echo "== ncu frontend rc=$?"
du -sh gpurun_out
