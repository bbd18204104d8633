set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 120 python tools/gpu_knn_bench.py 1000
timeout 120 python tools/gpu_knn_bench.py 1536
timeout 300 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_orb_match.py -m gpu -q -p no:cacheprovider -k "pyr or knn" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'fe_ms', d['roofline']['launch_ms'])"; tail -3 gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn2_partial -c 1 -f -o gpurun_out/prof_knn python tools/gpu_knn_bench.py 1000 > gpurun_out/ncu_knn.log 2>&1; echo "== ncu knn rc=$?"
