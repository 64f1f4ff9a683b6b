#!/bin/bash
# round 2, call i: register-resident TRSM version of the fused small-N kernel -- parity + A/B against version 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -12 gpurun_out/i_pytest.log
for v in new v1; do
  if [ $v = v1 ]; then export GPRN_SMALL_V1=1; else unset GPRN_SMALL_V1; fi
  timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu > gpurun_out/i_c3_$v.json 2> gpurun_out/i_c3_$v.err; echo "c3 $v rc=$?"
  timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu > gpurun_out/i_c2b_$v.json 2> gpurun_out/i_c2b_$v.err; echo "c2b $v rc=$?"
done
unset GPRN_SMALL_V1
timeout 300 python tools/trace_run.py 256 4 1 QP 4096 40 > gpurun_out/i_trace_c3.txt 2>&1; echo "trace rc=$?"
head -16 gpurun_out/i_trace_c3.txt
python - <<'PY'
import json
for f in ['i_c3_new','i_c3_v1','i_c2b_new','i_c2b_v1']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'e2e',d['e2e']['value'],'checksum',d['run']['elbo_checksum'],'meanit',d['run']['mean_iterations'])
    except Exception as e: print(f,'ERR',e)
PY
