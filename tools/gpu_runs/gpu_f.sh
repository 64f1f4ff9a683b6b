#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
timeout 700 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/f_c4_weak_n2.json 2> gpurun_out/f_c4_weak_n2.err; echo "weak n2 rc=$?"; tail -3 gpurun_out/f_c4_weak_n2.err
timeout 300 python tools/trace_run.py 256 4 1 QP 4096 5 > gpurun_out/f_trace_c3.txt 2>&1; echo "trace c3 rc=$?"; head -20 gpurun_out/f_trace_c3.txt
timeout 300 python tools/trace_run.py 4096 4 2 M52 8 3 > gpurun_out/f_trace_c4.txt 2>&1; echo "trace c4 rc=$?"; cat gpurun_out/f_trace_c4.txt | tail -32
