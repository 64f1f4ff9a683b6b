#!/bin/bash
mkdir -p gpurun_out
./tools/small_micro > gpurun_out/k_micro.txt 2>&1; cat gpurun_out/k_micro.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "factorisation or stress or batch_matches or c3_full or mid_n or elbo_matches" > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_pytest.log
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/k_c3_new.json 2> gpurun_out/k_c3_new.err; echo "c3 rc=$?"
GPRN_SMALL_ONE_PER_SM=1 timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/k_c3_one.json 2> gpurun_out/k_c3_one.err; echo "c3 one rc=$?"
timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/k_c2b_new.json 2> gpurun_out/k_c2b_new.err; echo "c2b rc=$?"
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/k_c2.json 2> gpurun_out/k_c2.err; echo "c2 rc=$?"
timeout 300 python tools/trace_run.py 256 4 1 QP 4096 40 > gpurun_out/k_trace_c3.txt 2>&1; echo "trace rc=$?"
head -14 gpurun_out/k_trace_c3.txt
python - <<'PY'
import json
for f in ['k_c3_new','k_c3_one','k_c2b_new','k_c2']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'checksum',d['run']['elbo_checksum'],'meanit',d['run']['mean_iterations'])
    except Exception as e: print(f,'ERR',e)
PY
