#!/usr/bin/env bash
# round 2: N = 2 (torchrun, NCCL): loop-closure exchange on the side stream; N = 1 on the same box for the efficiency
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n1_samebox.json 2> gpurun_out/bench_r2_n1_samebox.err
echo "== n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo "== n2 rc=$?"; tail -5 gpurun_out/bench_r2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-stage-stats --no-loop-closure > gpurun_out/bench_r2_n2_nolc.json 2> gpurun_out/bench_r2_n2_nolc.err
echo "== n2 (no loop closure) rc=$?"
python - <<'PY'
import json
def last(p):
    try: return json.loads([l for l in open(p).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e: return {"error": repr(e)}
a=last('gpurun_out/bench_r2_n1_samebox.json'); b=last('gpurun_out/bench_r2_n2.json'); c=last('gpurun_out/bench_r2_n2_nolc.json')
for n,d in (("n1",a),("n2",b),("n2 nolc",c)):
    print(n, {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, (d.get('e2e') or {}).get('value'), d.get('loop_closure'))
if 'value' in a and 'value' in b: print("efficiency N=2:", b['value']/(2*a['value']), " without exchange:", c.get('value',0)/(2*a['value']))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_ref_n2.json 2> gpurun_out/bench_r2_ref_n2.err
echo "== ref n2 rc=$?"; tail -2 gpurun_out/bench_r2_ref_n2.json | cut -c1-400
