set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for f in test_gpu_frontend test_gpu_orb_match test_gpu_pipeline test_gpu_system; do
  timeout 300 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "== $f rc=$?"; tail -5 gpurun_out/$f.log
done
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'fe_ms', d['roofline']['launch_ms'], d['clocks'], d['stats'])"; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ba-overlap > gpurun_out/ncu_bench.log 2>&1; echo "== ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frontend_tile -s 1 -c 1 -f -o gpurun_out/prof_frontend python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "== ncu full rc=$?"
