#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err ) 2> gpurun_out/n2_bench.time; echo "bench rc=$?"; tail -3 gpurun_out/n2_bench.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('sets per rank',d['run']['sets_per_rank_last_step'])
for k,v in d.get('also',{}).items(): print(k,'value',v['value'],'ms/step',v['ms_per_step'],'frac',v['roofline']['frac'],'e2e',v['e2e']['value'],'launches/step',v['gpu_launches']/v['steps'])
PY
tail -5 gpurun_out/n2_bench.err
