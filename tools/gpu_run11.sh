set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_frontend.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/gpu_step_breakdown.py 2>&1 | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:orb_describe -s 1 -c 1 -f -o gpurun_out/prof_describe python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_desc.log 2>&1; echo "== ncu describe rc=$?"
