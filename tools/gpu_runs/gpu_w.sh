#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mid or c3_full or batch_matches or smoke or continuous or elbo_matches" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/w_pytest.log
for v in new w4; do
  if [ $v = w4 ]; then export GPRN_MID_WARPS=4; else unset GPRN_MID_WARPS; fi
  timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/w_c2_$v.json 2> gpurun_out/w_c2_$v.err; echo "c2 $v rc=$?"
done
unset GPRN_MID_WARPS
python - <<'PY'
import json
for f in ['w_c2_new','w_c2_w4']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e']['value'],'launches/step',d['gpu_launches']/d['steps'],'graphs',d['graph_launches'],'checksum',d['run']['elbo_checksum'],'fail',d['run']['not_converged_or_failed'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 200 python tools/mid_phases.py > gpurun_out/w_phases.txt 2>&1; echo "phases rc=$?"; cat gpurun_out/w_phases.txt | tail -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_c2_launches.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/w_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/w_c2_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[hi+2:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[(r[ki].split('(')[0], r[gi])].append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(k, 'n',len(v),'mean us',round(sum(v)/len(v)/1e3,1),'share %',round(100*sum(v)/tot,1))
print('total us', tot/1e3)
PY
