#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/final2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final2_smoke.log
python tools/latency_probe.py 500 4 40 > gpurun_out/final2_latency.txt 2>&1; cat gpurun_out/final2_latency.txt
