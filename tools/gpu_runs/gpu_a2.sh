#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "elbo_matches or latency_path or batch_matches or more_than_two or continuous or per_set_means or stub_on" --durations=5 > gpurun_out/a2_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/a2_pytest.log
