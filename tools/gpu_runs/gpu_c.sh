#!/bin/bash
# round-2 GPU pass C (2 GPUs): full parity suite with the new potrf64, two-device test, N=2 weak scaling through torchrun
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --workload c4 --steps 2 --warmup 1 > gpurun_out/c_c4_n2.json 2> gpurun_out/c_c4_n2.err; echo "c4 n2 rc=$?"; tail -c 2000 gpurun_out/c_c4_n2.json; tail -3 gpurun_out/c_c4_n2.err
timeout 600 python bench.py --gpus 1 --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/c_c4_n1.json 2> gpurun_out/c_c4_n1.err; echo "c4 n1 rc=$?"; tail -c 600 gpurun_out/c_c4_n1.json
timeout 600 $TR bench.py --gpus 2 --impl reference --workload c4 --steps 2 --warmup 1 > gpurun_out/c_ref_n2.json 2> gpurun_out/c_ref_n2.err; echo "ref n2 rc=$?"; tail -c 1200 gpurun_out/c_ref_n2.json; tail -3 gpurun_out/c_ref_n2.err
timeout 300 $TR bench.py --gpus 2 --workload c3 --steps 2 --warmup 1 > gpurun_out/c_c3_n2.json 2> gpurun_out/c_c3_n2.err; echo "c3 n2 rc=$?"; tail -c 700 gpurun_out/c_c3_n2.json; tail -3 gpurun_out/c_c3_n2.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu > gpurun_out/c_c2.json 2> gpurun_out/c_c2.err; echo "c2 rc=$?"; tail -c 500 gpurun_out/c_c2.json
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/c_c3.json 2> gpurun_out/c_c3.err; echo "c3 rc=$?"; tail -c 500 gpurun_out/c_c3.json
