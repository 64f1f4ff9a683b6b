#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "mcmc or logposterior or optimize_batch or chain_state" > gpurun_out/i2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/i2_pytest.log
