#!/bin/bash
# round-2 GPU pass E (8 GPUs): default weak line, strong scaling on 128 natural-order sets, C5 sweep, reference arm
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 700 $TR bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/e_c4_weak_n8.json 2> gpurun_out/e_c4_weak_n8.err; echo "weak n8 rc=$?"; tail -c 400 gpurun_out/e_c4_weak_n8.json; tail -2 gpurun_out/e_c4_weak_n8.err
timeout 700 $TR bench.py --gpus 8 --scaling strong --pool 128 --slots 6 --steps 1 --warmup 1 --warmup-pool 12 --no-e2e --no-extra > gpurun_out/e_c4_strong_n8.json 2> gpurun_out/e_c4_strong_n8.err; echo "strong n8 rc=$?"; tail -c 1400 gpurun_out/e_c4_strong_n8.json | head -c 1000; tail -2 gpurun_out/e_c4_strong_n8.err
timeout 600 $TR bench.py --gpus 8 --workload c5 --steps 1 --maxfev 40 > gpurun_out/e_c5_n8.json 2> gpurun_out/e_c5_n8.err; echo "c5 n8 rc=$?"; tail -c 1200 gpurun_out/e_c5_n8.json; tail -2 gpurun_out/e_c5_n8.err
timeout 300 $TR bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/e_ref_n8.json 2> gpurun_out/e_ref_n8.err; echo "ref n8 rc=$?"; tail -c 300 gpurun_out/e_ref_n8.json
