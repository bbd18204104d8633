set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 python -m pytest tests/test_gpu_system.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_system.log 2>&1; echo "== system rc=$?"; tail -30 gpurun_out/test_gpu_system.log | cut -c1-300
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 --deselect tests/test_gpu_system.py > gpurun_out/test_gpu_all.log 2>&1; echo "== all rc=$?"; tail -30 gpurun_out/test_gpu_all.log | cut -c1-300
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_ref.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "== ncu list rc=$?"; wc -l gpurun_out/launches.csv
