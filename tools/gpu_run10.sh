set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 120 python tools/gpu_knn_bench.py 1000 8
timeout 120 python tools/gpu_knn_bench.py 1000 4
timeout 300 python -m pytest tests/test_gpu_orb_match.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'fe_ms', d['roofline']['launch_ms'])"; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ba-overlap > gpurun_out/ncu_bench.log 2>&1; echo "== ncu rc=$?"
