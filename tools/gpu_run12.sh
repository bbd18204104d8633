set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 200 python -m pytest tests/test_gpu_init.py -m gpu -x -q -p no:cacheprovider > gpurun_out/test_gpu_init.log 2>&1; echo "== init rc=$?"; tail -25 gpurun_out/test_gpu_init.log
timeout 200 python -m pytest tests/test_gpu_system.py -m gpu -x -q -p no:cacheprovider > gpurun_out/test_gpu_system.log 2>&1; echo "== system rc=$?"; tail -40 gpurun_out/test_gpu_system.log
