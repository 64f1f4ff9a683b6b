#!/bin/bash
# strong-scaling point: bash tools/gpu_runs/gpu_strong.sh N   (128 natural-order C4 sets, 6 in flight per GPU)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then CMD="python"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"; fi
timeout 1500 $CMD bench.py --gpus $N --scaling strong --pool 128 --slots 6 --steps 1 --warmup 1 --warmup-pool 12 --no-e2e --no-extra --no-cpu > gpurun_out/g_c4_strong_n$N.json 2> gpurun_out/g_c4_strong_n$N.err
echo "strong n$N rc=$?"; tail -c 600 gpurun_out/g_c4_strong_n$N.json; tail -2 gpurun_out/g_c4_strong_n$N.err
