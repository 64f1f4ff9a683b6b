#!/bin/bash
# round-2 GPU pass B: the tests pass A did not reach, kernel-level timing of the latency case, c5, full c4 line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 -k "continuous or work_source or chain_state or optimize_batch or logposterior or width_validation or ticket or two_devices or whitenoise_square or reference_integration or stub or 20000" > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_c2_launches.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu > gpurun_out/b_c2_ncu.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 python bench.py --workload c5 --steps 1 --maxfev 20 > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; echo "c5 rc=$?"; tail -c 1800 gpurun_out/b_c5.json; tail -5 gpurun_out/b_c5.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err; echo "c4 rc=$?"; tail -c 2500 gpurun_out/b_c4.json; tail -3 gpurun_out/b_c4.err
