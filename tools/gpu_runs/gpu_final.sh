#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
( time timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2> gpurun_out/final_bench.time; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],d['roofline'].get('frac_executed_min'),'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print('anchor',d['run']['anchor'])
print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'])
for k,v in d.get('also',{}).items(): print(k,'value',v['value'],'ms/step',v['ms_per_step'],'frac',v['roofline']['frac'],'e2e',v['e2e']['value'],'launches/step',v['gpu_launches']/v['steps'])
PY
