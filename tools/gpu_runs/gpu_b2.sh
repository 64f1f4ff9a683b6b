#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "latency_path or continuous or batch_matches or mid_n or chain_state or optimize_batch or logposterior" --durations=5 > gpurun_out/b2_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/b2_pytest.log
for v in new old; do
  if [ $v = old ]; then export GPRN_MID_COLOCATED_ONLY=1; else unset GPRN_MID_COLOCATED_ONLY; fi
  for pool in 16 32 128; do
    timeout 300 python bench.py --workload c2b --pool-per-gpu $pool --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/b2_c2b_${pool}_$v.json 2> gpurun_out/b2_c2b_${pool}_$v.err; echo "c2b $pool $v rc=$?"
  done
done
unset GPRN_MID_COLOCATED_ONLY
GPRN_NO_SMALL=1 timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/b2_c2b_512_midonly.json 2> gpurun_out/b2_c2b_512_midonly.err; echo "c2b 512 mid-only rc=$?"
timeout 300 python bench.py --workload c2b --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/b2_c2b_512_small.json 2> gpurun_out/b2_c2b_512_small.err; echo "c2b 512 small rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_c2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'frac',round(d['roofline']['frac'],3),'launches/step',d['gpu_launches']/d['steps'],'checksum',d['run']['elbo_checksum'],'fail',d['run']['not_converged_or_failed'])
    except Exception as e: print(f,'ERR',e)
PY
