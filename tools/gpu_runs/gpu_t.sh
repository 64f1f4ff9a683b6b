#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/t_pytest.log
for v in mid nomid; do
  if [ $v = nomid ]; then export GPRN_NO_MID=1; else unset GPRN_NO_MID; fi
  timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/t_c2_$v.json 2> gpurun_out/t_c2_$v.err; echo "c2 $v rc=$?"
done
unset GPRN_NO_MID
python - <<'PY'
import json
for f in ['t_c2_mid','t_c2_nomid']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',d['e2e']['value'],'launches/step',d['gpu_launches']/d['steps'],'graphs',d['graph_launches'],'checksum',d['run']['elbo_checksum'],'fail',d['run']['not_converged_or_failed'])
    except Exception as e: print(f,'ERR',e)
PY
