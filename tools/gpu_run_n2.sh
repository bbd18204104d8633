set +e
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_orb_match.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_orb_match.log 2>&1; echo "== orb rc=$?"; tail -8 gpurun_out/test_gpu_orb_match.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "== bench n2 rc=$?"; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-loop-closure > gpurun_out/bench_n2_nolc.json 2> gpurun_out/bench_n2_nolc.err; echo "== bench n2 nolc rc=$?"; tail -3 gpurun_out/bench_n2_nolc.err; cat gpurun_out/bench_n2_nolc.json
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "== ref n2 rc=$?"; tail -3 gpurun_out/bench_ref_n2.err; cat gpurun_out/bench_ref_n2.json
NCCL_DEBUG=INFO timeout 200 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep -i "nvls\|NVLink\|P2P\|via" | head -12
