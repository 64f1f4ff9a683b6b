#!/bin/bash
mkdir -p gpurun_out
( time GPRN_NO_GRAPH=1 timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/h2_launches_bench_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra > gpurun_out/h2_ncu.log 2>&1 ) 2> gpurun_out/h2_time.txt; echo "ncu rc=$?"; tail -3 gpurun_out/h2_time.txt
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/h2_launches_bench_c4.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hi+2:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    name=r[ki].split('(')[0].replace('void ','').replace('gprn::','')
    agg[name].append(v)
tot=sum(sum(v) for v in agg.values()); n=sum(len(v) for v in agg.values())
out=[]
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    out.append(f"{k:40s} n={len(v):5d} total {sum(v)/1e6:9.3f} ms  avg {sum(v)/len(v)/1e3:9.1f} us  {100*sum(v)/tot:5.1f}%")
out.append(f"total {tot/1e6:.1f} ms over {n} launches")
open('gpurun_out/h2_launches_summary.txt','w').write("\n".join(out)+"\n")
print("\n".join(out[:16])); print(out[-1])
PY
