#!/bin/bash
# round-2 GPU pass A: parity tests, then first bench lines (c4 with two slot counts, c3 v2/v1 potrf, c2)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
timeout 400 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/a_c4_all.json 2> gpurun_out/a_c4_all.err; echo "c4 all rc=$?"; tail -c 1500 gpurun_out/a_c4_all.json
timeout 400 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu --slots 4 > gpurun_out/a_c4_s4.json 2> gpurun_out/a_c4_s4.err; echo "c4 s4 rc=$?"; tail -c 600 gpurun_out/a_c4_s4.json
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/a_c3.json 2> gpurun_out/a_c3.err; echo "c3 rc=$?"; tail -c 900 gpurun_out/a_c3.json
GPRN_B200_LIB=$PWD/gpyrn_b200/csrc/libgprn_b200_v1.so timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/a_c3_v1.json 2> gpurun_out/a_c3_v1.err; echo "c3 v1 rc=$?"; tail -c 400 gpurun_out/a_c3_v1.json
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu > gpurun_out/a_c2.json 2> gpurun_out/a_c2.err; echo "c2 rc=$?"; tail -c 900 gpurun_out/a_c2.json
GPRN_B200_LIB=$PWD/gpyrn_b200/csrc/libgprn_b200_v1.so timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu > gpurun_out/a_c2_v1.json 2> gpurun_out/a_c2_v1.err; echo "c2 v1 rc=$?"; tail -c 300 gpurun_out/a_c2_v1.json
tail -5 gpurun_out/*.err
