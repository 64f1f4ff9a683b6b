#!/bin/bash
mkdir -p gpurun_out
GPRN_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mid_ -c 120 --csv --log-file gpurun_out/v_mid_launches.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/v_mid_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[hi+2:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[(r[ki].split('(')[0], r[gi])].append(v)
for k,v in sorted(agg.items()):
    print(k, 'n',len(v),'mean us',round(sum(v)/len(v)/1e3,1),'min',round(min(v)/1e3,1),'max',round(max(v)/1e3,1))
PY
