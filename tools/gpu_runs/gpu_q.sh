#!/bin/bash
# round 2, call q: full GPU suite on the cleaned build + the named small configurations + the C3 phase trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/q_smoke.log
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/q_c3.json 2> gpurun_out/q_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --workload c2b --steps 3 --warmup 2 --no-cpu > gpurun_out/q_c2b.json 2> gpurun_out/q_c2b.err; echo "c2b rc=$?"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/q_c2.json 2> gpurun_out/q_c2.err; echo "c2 rc=$?"
timeout 300 python tools/trace_run.py 256 4 1 QP 4096 40 > gpurun_out/q_trace_c3.txt 2>&1; echo "trace rc=$?"
python - <<'PY'
import json
for f in ['q_c3','q_c2b','q_c2']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'meanit',d['run']['mean_iterations'],'cpu',d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
