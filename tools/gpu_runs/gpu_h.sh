#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -14 gpurun_out/h_pytest.log
timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/h_c2b_small.json 2> gpurun_out/h_c2b_small.err; echo "c2b small rc=$?"; tail -2 gpurun_out/h_c2b_small.err
GPRN_SMALL_MAX_NT=4 timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/h_c2b_big.json 2> gpurun_out/h_c2b_big.err; echo "c2b big rc=$?"; tail -2 gpurun_out/h_c2b_big.err
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/h_c4.json 2> gpurun_out/h_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/h_c4.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/h_smoke.log
python - <<'PY'
import json
for f in ['h_c2b_small','h_c2b_big','h_c4']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],3),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'meanit',d['run']['mean_iterations'])
    except Exception as e: print(f,'ERR',e)
PY
