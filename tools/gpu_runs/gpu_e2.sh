#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "up_to_n_1024 or continuous or latency_path_with" --durations=4 > gpurun_out/e2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/e2_pytest.log
{
python tools/latency_probe.py 1000 4 10
GPRN_NO_MID=1 python tools/latency_probe.py 1000 4 5
python tools/latency_probe.py 700 4 10
GPRN_NO_MID=1 python tools/latency_probe.py 700 4 5
python tools/latency_probe.py 1000 1 10
GPRN_NO_MID=1 python tools/latency_probe.py 1000 1 5
} > gpurun_out/e2_latency.txt 2>&1
cat gpurun_out/e2_latency.txt
timeout 200 python tools/mid_phases.py 1000 4 > gpurun_out/e2_phases.txt 2>&1; echo "phases rc=$?"; tail -20 gpurun_out/e2_phases.txt
