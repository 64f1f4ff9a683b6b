#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "kernel_matrices or elbo_matches or prediction_matches or stub_on or continuous or c3_full" > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/z_pytest.log
{
python tools/latency_probe.py 500 4 60
GPRN_NO_LOOP=1 python tools/latency_probe.py 500 4 60
python tools/latency_probe.py 497 1 60
} > gpurun_out/z_latency.txt 2>&1
cat gpurun_out/z_latency.txt
