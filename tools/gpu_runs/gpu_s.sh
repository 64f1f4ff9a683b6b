#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "batch_matches or c3_full or mid_n or elbo_matches or continuous or warm or chain" > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s_pytest.log
for v in tma notma tma2 notma2; do
  if [ ${v:0:5} = notma ]; then export GPRN_SMALL_NO_TMA=1; else unset GPRN_SMALL_NO_TMA; fi
  timeout 200 python bench.py --workload c3 --steps 4 --warmup 2 --no-cpu > gpurun_out/s_c3_$v.json 2> gpurun_out/s_c3_$v.err; echo "c3 $v rc=$?"
done
for v in tma notma; do
  if [ $v = notma ]; then export GPRN_SMALL_NO_TMA=1; else unset GPRN_SMALL_NO_TMA; fi
  timeout 200 python bench.py --workload c2b --steps 3 --warmup 2 --no-cpu > gpurun_out/s_c2b_$v.json 2> gpurun_out/s_c2b_$v.err; echo "c2b $v rc=$?"
done
python - <<'PY'
import json
for f in ['s_c3_tma','s_c3_notma','s_c3_tma2','s_c3_notma2','s_c2b_tma','s_c2b_notma']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],4),'checksum',d['run']['elbo_checksum'],'fail',d['run']['not_converged_or_failed'])
    except Exception as e: print(f,'ERR',e)
PY
