#!/usr/bin/env bash
# round 2, last visit: the whole GPU suite on the final tree, smoke, the bench lines again
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -4 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "== bench rc=$?"; tail -3 gpurun_out/r02_bench.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-stage-stats > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
echo "== bench c3 rc=$?"
python - <<'PY'
import json
for f in ('r02_bench','r02_bench_c3'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','output_check','gpu_launches')}, (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
