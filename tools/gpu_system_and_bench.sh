set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_system.py tests/test_gpu_init.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_system.log 2>&1; echo "== system+init rc=$?"; tail -30 gpurun_out/test_gpu_system.log | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -c 1800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
