#!/bin/bash
mkdir -p gpurun_out
{
python tools/latency_probe.py 500 4 60
GPRN_B200_LIB=$PWD/gpurun_in/libgprn_prev.so python tools/latency_probe.py 500 4 60
python tools/latency_probe.py 497 1 20
GPRN_B200_LIB=$PWD/gpurun_in/libgprn_prev.so python tools/latency_probe.py 497 1 20
} > gpurun_out/g2_latency.txt 2>&1
cat gpurun_out/g2_latency.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "c3_full or mid_n or up_to_n_1024 or elbo_matches_reference" > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g2_pytest.log
