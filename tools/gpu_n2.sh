#!/usr/bin/env bash
# round 2: N = 2 again (geometric check on the newest keyframe only) + the loop-closure tests
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests/test_gpu_loopclosure.py -x -q 2>&1 | grep -E "^E |assert|passed|failed" | head -20
python tools/gpu_lc_bench.py 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n1_samebox.json 2> gpurun_out/bench_r2_n1_samebox.err
echo "== n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo "== n2 rc=$?"; tail -5 gpurun_out/bench_r2_n2.err | cut -c1-300
python - <<'PY'
import json
def last(p):
    try: return json.loads([l for l in open(p).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e: return {"error": repr(e)}
a=last('gpurun_out/bench_r2_n1_samebox.json'); b=last('gpurun_out/bench_r2_n2.json')
for n,d in (("n1",a),("n2",b)):
    print(n, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'), d.get('loop_closure'))
if 'value' in a and 'value' in b: print("efficiency N=2:", b['value']/(2*a['value']))
PY
