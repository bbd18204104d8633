#!/usr/bin/env bash
# One GPU-box visit: parity tests file by file (a crashed CUDA context only poisons its own process), a short bench,
# the ncu launch list of the bench command and one full capture of the dominant kernel.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
FILES=${FILES:-"test_gpu_frontend test_gpu_orb_match test_gpu_ba test_gpu_pipeline test_gpu_klt test_gpu_pose test_gpu_detect test_gpu_match test_gpu_init test_gpu_system test_gpu_zz_experimental"}
rm -f gpurun_out/test_gpu_*.log
for f in $FILES; do
  timeout 420 python -m pytest tests/$f.py -m gpu -x -q --durations=4 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f rc=$?"; tail -14 gpurun_out/$f.log
done
if [ -f gpurun_out/test_gpu_frontend.log ] && { ! grep -q " passed" gpurun_out/test_gpu_frontend.log || grep -q "failed" gpurun_out/test_gpu_frontend.log; }; then
  echo "== frontend failed: retry without TMA and under compute-sanitizer"
  ALVA_DISABLE_TMA=1 timeout 300 python -m pytest tests/test_gpu_frontend.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
  timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_repro.py > gpurun_out/sanitizer.log 2>&1; tail -30 gpurun_out/sanitizer.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 400 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "== bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -4 gpurun_out/bench.err
  timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "== ncu list rc=$?"; wc -l gpurun_out/launches.csv
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:frontend_tile -s 1 -c 2 -f -o gpurun_out/prof_frontend \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "== ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
# experimental front-end instantiation (antipodal flag sharing + hoisted gray-phase tests): A/B timing of the fused launch
if [ "${SKIP_AB:-0}" != "1" ]; then
  timeout 200 python - <<'PY' 2>&1 | tail -4
import torch, numpy as np, alvaar_b200
from alvaar_b200 import synth
ctx = alvaar_b200.Context(0, torch.cuda.current_stream().cuda_stream)
n, w, h = 64, 1280, 720
fr, _ = synth.make_frames(4, w, h)
d = torch.from_numpy(np.ascontiguousarray(np.tile(fr, (16, 1, 1, 1)))).cuda()
l0 = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda"); l1 = torch.zeros((n, h // 2, w // 2), dtype=torch.uint8, device="cuda")
keys = torch.zeros((n, 32768), dtype=torch.int32, device="cuda"); cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
for opt in (0, 1, 0, 1):
    ctx.L.alva_set_option(b"frontend_antipodal", opt)
    ts = []
    for _ in range(8):
        cnt.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.frontend(d, w, h, n, l0, l1, None, None, 20, keys, cnt, 32768, False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("frontend_antipodal =", opt, "median us (front end + level 2/3 + order):", round(float(np.median(ts[2:])), 1), "corners", int(cnt.sum()))
ctx.L.alva_set_option(b"frontend_antipodal", 0)
PY
fi
