set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:frontend_tile -s 1 -c 1 -f -o gpurun_out/prof_frontend_r01f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "== ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full.log
