#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "batch_matches or c3_full or mid_n or elbo_matches or continuous or warm or chain" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/n_pytest.log
timeout 200 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/n_c3_v4.json 2> gpurun_out/n_c3_v4.err; echo "c3 v4 rc=$?"
GPRN_SMALL_CTAS=1 timeout 200 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/n_c3_v4_one.json 2> gpurun_out/n_c3_v4_one.err; echo "c3 v4 one rc=$?"
timeout 200 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/n_c2b_v4.json 2> gpurun_out/n_c2b_v4.err; echo "c2b rc=$?"
timeout 200 python tools/trace_run.py 256 4 1 QP 4096 40 > gpurun_out/n_trace_c3.txt 2>&1; echo "trace rc=$?"
head -14 gpurun_out/n_trace_c3.txt
python - <<'PY'
import json
for f in ['n_c3_v4','n_c3_v4_one','n_c2b_v4']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'checksum',d['run']['elbo_checksum'],'meanit',d['run']['mean_iterations'],'fail',d['run']['not_converged_or_failed'])
    except Exception as e: print(f,'ERR',e)
PY
