#!/bin/bash
# round-2 GPU job 24: per-kernel split of top-k recommend
mkdir -p gpurun_out
timeout 300 python tools/bench_recommend.py > gpurun_out/r2_recommend24.json 2> gpurun_out/r2_recommend24.err
cat gpurun_out/r2_recommend24.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_recommend.csv \
    python tools/bench_recommend.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_launches_recommend.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
    if len(r)<=iv: continue
    v=float(r[iv].replace(',','')); u=r[iu]
    ms=v/1e6 if u.startswith('n') else v/1e3 if u.startswith('u') else v
    name=r[ik].split('(')[0][-50:]
    tot[name]+=ms; cnt[name]+=1
for k,v in tot.most_common(8): print("%-52s %3d launches %9.3f ms"%(k,cnt[k],v))
PY
echo job24 done
