#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mid_n or latency_path or c3_full" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c2_pytest.log
timeout 300 python bench.py --workload c2b --pool-per-gpu 128 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/c2_c2b_128.json 2> gpurun_out/c2_c2b_128.err; echo "c2b 128 rc=$?"
timeout 300 python bench.py --workload c2b --pool-per-gpu 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/c2_c2b_256.json 2> gpurun_out/c2_c2b_256.err; echo "c2b 256 rc=$?"
GPRN_SMALL_MIN_FILL=2 timeout 300 python bench.py --workload c2b --pool-per-gpu 256 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/c2_c2b_256_small.json 2> gpurun_out/c2_c2b_256_small.err; echo "c2b 256 small rc=$?"
GPRN_NO_SMALL=1 timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/c2_c3_midonly.json 2> gpurun_out/c2_c3_midonly.err; echo "c3 mid-only rc=$?"
timeout 300 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu > gpurun_out/c2_c2.json 2> gpurun_out/c2_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c2_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'launches/step',d['gpu_launches']/d['steps'],'e2e',d.get('e2e',{}).get('value'),'checksum',d['run']['elbo_checksum'],'fail',d['run']['not_converged_or_failed'], d.get('clocks'))
    except Exception as e: print(f,'ERR',e)
PY
{
python tools/latency_probe.py 500 4 60
python tools/latency_probe.py 497 1 30
python tools/latency_probe.py 256 4 40 2 M52
GPRN_NO_MID=1 python tools/latency_probe.py 256 4 10 2 M52
python tools/latency_probe.py 100 4 40 2 M52
GPRN_NO_MID=1 python tools/latency_probe.py 100 4 10 2 M52
} > gpurun_out/c2_latency.txt 2>&1
cat gpurun_out/c2_latency.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mid_pipeline -s 20 -c 2 -o gpurun_out/c2_mid_full -f python tools/latency_probe.py 500 4 3 > gpurun_out/c2_ncu_full.log 2>&1; echo "ncu full rc=$?"
