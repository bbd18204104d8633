set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for f in test_gpu_pipeline test_gpu_system; do
  timeout 300 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider > gpurun_out/$f.log 2>&1; echo "== $f rc=$?"; tail -4 gpurun_out/$f.log
done
for opt in "" "--no-ba-overlap"; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $opt > gpurun_out/bench$opt.json 2> gpurun_out/bench.err; echo "== bench $opt rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench$opt.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'fe_ms', d['roofline']['launch_ms'], d['clocks'])"; tail -3 gpurun_out/bench.err
done
