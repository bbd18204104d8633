set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -c 700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
