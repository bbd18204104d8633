set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_system.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_system.log 2>&1; echo "== system rc=$?"; tail -40 gpurun_out/test_gpu_system.log | cut -c1-400
