#!/usr/bin/env bash
# round 2: N = 8 (and 4) with the loop-closure exchange; N = 1 on the same box for the efficiency
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n1_box8.json 2> gpurun_out/bench_r2_n1_box8.err
echo "== n1 rc=$?"
for N in 8; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "== n$N rc=$?"; grep -v 'OMP_NUM_THREADS\|^\*\*\*\|^$' gpurun_out/bench_r2_n$N.err | tail -4 | cut -c1-300
done
python - <<'PY'
import json
def last(p):
    try: return json.loads([l for l in open(p).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e: return {"error": repr(e)}
a=last('gpurun_out/bench_r2_n1_box8.json')
print("n1", {k:a.get(k) for k in ('value','ms_per_step')})
for N in (8,):
    b=last(f'gpurun_out/bench_r2_n{N}.json')
    print(f"n{N}", {k:b.get(k) for k in ('value','ms_per_step','n_gpus')}, (b.get('e2e') or {}).get('value'), b.get('loop_closure'))
    if 'value' in a and 'value' in b: print(f"efficiency N={N}:", b['value']/(N*a['value']))
PY
