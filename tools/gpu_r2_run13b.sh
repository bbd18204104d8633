#!/usr/bin/env bash
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python tools/gpu_lc_bench.py 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lc.csv python tools/gpu_lc_bench.py > gpurun_out/ncu_lc.log 2>&1
python - <<'PY'
import csv,re,collections
rows=list(csv.reader(open('gpurun_out/launches_lc.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value")
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    n=re.sub(r'\(.*','',r[kn]).replace('<unnamed>::','').replace('void ','')
    tot[n]+=float(r[mv])/1000; cnt[n]+=1
for n,t in tot.most_common(14): print(f"{n:40s} n={cnt[n]:3d} total {t:9.1f} us  avg {t/cnt[n]:8.1f}")
PY
bash tools/gpu_r2_run13.sh
