#!/bin/bash
# round 2, 8 GPUs: the driver's weak-scaling command (fewer steps) after the fair-share cap on sets in flight per rank
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 500 $TR bench.py --gpus 8 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/r_c4_weak_n8.json 2> gpurun_out/r_c4_weak_n8.err; echo "weak n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r_c4_weak_n8.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
print('sets per rank',d['run']['sets_per_rank_last_step'],'iters',d['run']['iterations_per_rank_last_step'])
for k,v in d.get('also',{}).items(): print(k,'value',v['value'],'ms/step',v['ms_per_step'],'frac',v['roofline']['frac'],'e2e',v['e2e']['value'])
PY
tail -3 gpurun_out/r_c4_weak_n8.err
