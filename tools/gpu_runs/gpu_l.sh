#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "factorisation or stress or batch_matches or c3_full or mid_n or elbo_matches or continuous or warm or chain or width" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/l_pytest.log
for n in 4 3 2; do
GPRN_SMALL_CTAS=$n timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/l_c3_n$n.json 2> gpurun_out/l_c3_n$n.err; echo "c3 ctas=$n rc=$?"
done
GPRN_SMALL_V2=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/l_c3_v2.json 2> gpurun_out/l_c3_v2.err; echo "c3 v2 rc=$?"
timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/l_c2b_n4.json 2> gpurun_out/l_c2b_n4.err; echo "c2b rc=$?"
GPRN_SMALL_CTAS=3 timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/l_c2b_n3.json 2> gpurun_out/l_c2b_n3.err; echo "c2b n3 rc=$?"
timeout 300 python tools/trace_run.py 256 4 1 QP 4096 40 > gpurun_out/l_trace_c3.txt 2>&1; echo "trace rc=$?"
head -14 gpurun_out/l_trace_c3.txt
python - <<'PY'
import json
for f in ['l_c3_n4','l_c3_n3','l_c3_n2','l_c3_v2','l_c2b_n4','l_c2b_n3']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'checksum',d['run']['elbo_checksum'],'meanit',d['run']['mean_iterations'])
    except Exception as e: print(f,'ERR',e)
PY
