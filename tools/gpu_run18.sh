set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "== bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err; tail -c 900 gpurun_out/bench_n2.json
